"""Two-field (mixed u-p) assembly on the device: MultiFieldCellValues + the element routine of the
incompressible-elasticity tutorial, against oracle/mixed.py (pinned on the tutorial's literal in tests/test_oracle_goldens.py)
and against that literal itself."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import ferrite_b200 as fb
import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return fb.default_context(0)


def close(a, b, tol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all(np.abs(a - b) <= tol * max(np.abs(b).max(), 1e-300)))


def cook(n):
    g = fb.generate_grid(fb.Triangle, (n, n), corners=[(0.0, 0.0), (48.0, 44.0), (48.0, 60.0), (0.0, 44.0)])
    fb.addfacetset_(g, "clamped", lambda x: abs(x[0]) <= 1e-8)
    fb.addfacetset_(g, "traction", lambda x: abs(x[0] - 48.0) <= 1e-8)
    return g


@pytest.mark.parametrize("n,order_u", [(6, 2), (5, 1)])
def test_mixed_up_matches_oracle(ctx, n, order_u):
    g, og = cook(n), O.cook_grid(n, n)
    assert np.allclose(g.nodes, og.nodes, rtol=0, atol=1e-13) and np.array_equal(g.cells, og.cells)
    assert np.array_equal(fb.getfacetset(g, "traction"), og.facetsets["traction"])
    ipu, ipp = fb.Lagrange(fb.RefTriangle, order_u) ** 2, fb.Lagrange(fb.RefTriangle, 1)
    oipu, oipp = O.Lagrange("triangle", order_u) ** 2, O.Lagrange("triangle", 1)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", ipu), "p", ipp))
    odh = O.DofHandler(og).add("u", oipu).add("p", oipp).close()
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    K, oK = fb.allocate_matrix(dh), O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    cvs = fb.MultiFieldCellValues(fb.QuadratureRule(fb.RefTriangle, 3), u=ipu, p=ipp)
    oqr = O.QuadratureRule("triangle", 3)
    G, Kb = 0.37, 2.5
    f = ctx.zeros(dh.ndofs)
    fb.assemble_mixed_up_(fb.start_assemble(K, f), cvs, G, Kb)
    O.assemble_up(odh, O.CellValues(oqr, oipu), O.CellValues(oqr, oipp), oK, G, 1.0 / Kb)
    assert close(K.nzval.cpu().numpy(), oK.nzval)
    # two different quadrature rules are not a MultiFieldCellValues
    bad = fb.MultiFieldCellValues(fb.QuadratureRule(fb.RefTriangle, 3), u=ipu, p=ipp)
    bad.p = fb.CellValues(fb.QuadratureRule(fb.RefTriangle, 2), ipp)
    with pytest.raises(fb.FB2Error, match="share the quadrature rule"):
        fb.assemble_mixed_up_(fb.start_assemble(K, f), bad, G, Kb)


def test_incompressible_elasticity_tutorial_golden_through_the_gpu_path(ctx):
    """docs/src/literate-tutorials/incompressible_elasticity.jl:386-446, quadratic / linear triangles on Cook's membrane 50 x 50,
    nu = 0.5 (1 / K = 0): mixed assembly, traction facet term and apply! on the device, norm(u) = 919.1284143115702 (:477)"""
    g = cook(50)
    ipu, ipp = fb.Lagrange(fb.RefTriangle, 2) ** 2, fb.Lagrange(fb.RefTriangle, 1)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", ipu), "p", ipp))
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    cvs = fb.MultiFieldCellValues(fb.QuadratureRule(fb.RefTriangle, 3), u=ipu, p=ipp)
    fb.assemble_mixed_up_(fb.start_assemble(K, f), cvs, 1.0 / (2 * 1.5), np.inf)
    fv = fb.FacetValues(fb.FacetQuadratureRule(fb.RefTriangle, 3), ipu)
    fb.assemble_facets_(f, dh, fv, fb.getfacetset(g, "traction"), "traction", (0.0, 1.0 / 16.0))
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "clamped"), lambda x, t: [0.0, 0.0], [1, 2]))
    fb.close_(ch)
    fb.update_(ch, 0.0)
    fb.apply_(K, f, ch)
    u = spla.spsolve(K.tocsc(), f.cpu().numpy())
    assert abs(np.linalg.norm(u) - 919.1284143115702) <= 1e-8 * 919.1284143115702

"""GPU tests of the matrix-free solver path: CG on the element-assembly operator, pinned on the heat tutorial's
norm(u) golden, and the Poisson convergence rates of test/integration/convergence_test_utils.jl computed entirely on the device.
(Sorted after the parity files on purpose: these are end-to-end physics checks built on top of them.)"""
import numpy as np
import pytest

import ferrite_b200 as fb
import oracle as O
from test_gpu_element_assembly import dirichlet
from test_gpu_parity import build, close, make_element

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return fb.default_context(0)


def test_matrix_free_cg_solves_the_heat_tutorial(ctx):
    # heat_equation.jl:59-114,181-234 without a global matrix (the solver loop of gpu_assembly.jl:287-304): element matrices,
    # apply_local!, CG on y = sum_e P' Ke P x; norm(u) == 3.307743912641305
    g = fb.generate_grid(fb.Quadrilateral, (20, 20))
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip)
    ch = fb.ConstraintHandler(dh)
    boundary = np.concatenate([fb.getfacetset(g, k) for k in ("left", "right", "top", "bottom")])
    fb.add_(ch, fb.Dirichlet("u", boundary, lambda x, t: 0))
    fb.close_(ch)
    ea = fb.ElementAssembly(dh, cv)
    Kes, fes = ea.assemble(fb.HeatElement())
    ea.apply_local_(Kes, fes, ch)
    f = ea.rhs(fes)
    ref = 3.307743912641305
    for jacobi in (False, True):
        u = ctx.zeros(dh.ndofs)
        it, rn = ea.cg_(u, Kes, f, reltol=1e-13, jacobi=jacobi)
        assert 0 < it < dh.ndofs and rn <= 1e-13 * float(f.norm()) * 1.01
        assert abs(float(u.norm()) - ref) / ref < 1e-11
    # rhs / diag against the assembled system of the same element matrices
    K = fb.allocate_matrix(dh)
    fa = ctx.zeros(dh.ndofs)
    fb.scatter_device_(fb.start_assemble(K, fa), Kes, fes)
    assert close(f.cpu().numpy(), fa.cpu().numpy())[0]
    assert close(ea.diag(Kes).cpu().numpy(), K.tocsc().diagonal())[0]


def test_matrix_free_cg_elasticity_matches_assembled_solve(ctx):
    # Q1^3 elasticity with inhomogeneous Dirichlet values: matrix-free CG == sparse direct solve of apply_assemble!'s system
    import scipy.sparse.linalg as spla
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (5, 4, 3), 1, 3, 2)
    elem, op = make_element("elasticity", {"E": 10.0, "nu": 0.3, "b": (0.0, 0.0, -1.0)})
    ch, och = dirichlet(g, og, dh, odh, 3)
    ea = fb.ElementAssembly(dh, cv)
    Kes, fes = ea.assemble(elem)
    ea.apply_local_(Kes, fes, ch)
    f = ea.rhs(fes)
    u = ctx.zeros(dh.ndofs)
    it, rn = ea.cg_(u, Kes, f, reltol=1e-13, jacobi=True)
    assert 0 < it < dh.ndofs
    K = fb.allocate_matrix(dh)
    fa = ctx.zeros(dh.ndofs)
    fb.apply_assemble_(fb.start_assemble(K, fa), ch, elem, cv, ea=ea)
    uref = spla.spsolve(K.tocsc(), fa.cpu().numpy())
    assert close(u.cpu().numpy(), uref, 1e-9)[0]
    pd = och.prescribed_dofs - 1
    assert np.allclose(u.cpu().numpy()[pd], och.inhomogeneities, rtol=1e-10, atol=1e-13)


def _poisson_errors_gpu(ctx, ct, order, N, matrix_free):
    # the GPU twin of tests/test_oracle_goldens.py::_poisson_errors_oracle (test/integration/convergence_test_utils.jl):
    # heat + mass assembly, apply! / apply_local!, CG and function_value / function_gradient all on the device
    import torch
    dim = 2 if ct == fb.Quadrilateral else 3
    shape = "quadrilateral" if dim == 2 else "hexahedron"
    g = fb.generate_grid(ct, (N,) * dim)
    ip = fb.Lagrange(ct, order)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(ct, max(2 * order - 1, 2)), ip)
    ana = lambda x: np.prod(np.cos(np.pi * np.asarray(x) / 2), axis=-1)     # noqa: E731
    ch = fb.ConstraintHandler(dh)
    boundary = np.concatenate([fb.getfacetset(g, k) for k in ("left", "right", "top", "bottom") + (("front", "back") if dim == 3 else ())])
    fb.add_(ch, fb.Dirichlet("u", boundary, lambda x, t: float(ana(x))))
    fb.close_(ch)
    fb.update_(ch, 0.0)
    oip, geo = O.Lagrange(shape, order), O.Lagrange(shape, 1)
    cd = dh.cell_dofs
    xn = g.nodes[g.cells - 1]
    xd = np.zeros((dh.ndofs, dim))
    for a in range(oip.nbase):
        M, _ = geo.value_and_gradient(oip.refcoords[a])
        xd[cd[:, a] - 1] = np.einsum("j,cjd->cd", M, xn)
    dev = f"cuda:{ctx.device}"
    rhs = torch.from_numpy(dim * np.pi ** 2 / 4 * ana(xd)).to(dev)
    u = ctx.zeros(dh.ndofs)
    if matrix_free:
        # no global matrix anywhere: fe = Me rhs_e per cell, apply_local!, CG on the element-assembly operator
        ea = fb.ElementAssembly(dh, cv)
        Mes, _ = ea.assemble(fb.MassElement(1.0))
        Kes, fes = ea.assemble(fb.HeatElement(1.0, 0.0))
        rhs_e = rhs[torch.from_numpy(cd - 1).to(dev)]
        fes.copy_(torch.einsum("cji,cj->ci", Mes, rhs_e))
        ea.apply_local_(Kes, fes, ch)
        it, _ = ea.cg_(u, Kes, ea.rhs(fes), reltol=1e-12, jacobi=True)
    else:
        Mm = fb.allocate_matrix(dh)
        fb.assemble_(fb.start_assemble(Mm, None), fb.MassElement(1.0), cv)
        f = fb.spmv(Mm, rhs)
        K = fb.allocate_matrix(dh)
        fb.assemble_(fb.start_assemble(K, None), fb.HeatElement(1.0, 0.0), cv)
        fb.apply_(K, f, ch)
        it, _ = fb.cg_(u, K, f, reltol=1e-12, jacobi=True)
    assert 0 < it < dh.ndofs
    fb.apply_(u, ch)
    uh, gh = fb.function_values_(cv, dh, u)
    _, dO = fb.reinit_(cv, g)
    xq = torch.stack([fb.function_values_(cv, dh, torch.from_numpy(np.ascontiguousarray(xd[:, k])).to(dev), gradients=False)[..., 0]
                      for k in range(dim)], dim=-1).cpu().numpy()
    uh, gh, dO = uh.cpu().numpy()[..., 0], gh.cpu().numpy()[..., 0, :], dO.cpu().numpy()
    ua = ana(xq)
    ga = np.stack([-np.pi / 2 * np.tan(np.pi * xq[..., k] / 2) * ua for k in range(dim)], axis=-1)
    return (np.sqrt(np.sum((ua - uh) ** 2 * dO)), np.sqrt(np.sum(np.sum((ga - gh) ** 2, axis=-1) * dO)), np.abs(ua - uh).max())


@pytest.mark.parametrize("ct,order,N", [(fb.Quadrilateral, 1, 21), (fb.Quadrilateral, 2, 7), (fb.Hexahedron, 1, 11)])
@pytest.mark.parametrize("matrix_free", [False, True])
def test_poisson_convergence_rates_on_the_device(ctx, ct, order, N, matrix_free):
    # L2 rate = order + 1, H1 rate = order, atol 0.1 (test/integration/convergence_test_utils.jl:176-207)
    l1, h1, m1 = _poisson_errors_gpu(ctx, ct, order, N, matrix_free)
    l2, h2, m2 = _poisson_errors_gpu(ctx, ct, order, 2 * N, matrix_free)
    assert m1 < 3e-2 and m2 < 1e-2
    assert abs(np.log(l1 / l2) / np.log(2) - (order + 1)) < 0.1
    assert abs(np.log(h1 / h2) / np.log(2) - order) < 0.1


@pytest.mark.parametrize("applyzero", [False, True])
def test_apply_rhs_matches_apply_and_oracle(ctx, applyzero):
    # test/test_apply_rhs.jl:5-92 on the device, and parity of the right-hand side with the oracle's apply_rhs!
    g, og, dh, odh, cv, ocv = build(fb.Quadrilateral, (20, 20), 1, 1, 2)
    elem, op = make_element("heat", {})
    ch, och = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    lr = ("left", "right")
    tb = ("top", "bottom")
    fb.add_(ch, fb.Dirichlet("u", np.concatenate([fb.getfacetset(g, k) for k in lr]), lambda x, t: 0))
    fb.add_(ch, fb.Dirichlet("u", np.concatenate([fb.getfacetset(g, k) for k in tb]), lambda x, t: 2 + t * x[0]))
    och.add(O.Dirichlet("u", np.concatenate([og.facetsets[k] for k in lr]), lambda x, t: 0))
    och.add(O.Dirichlet("u", np.concatenate([og.facetsets[k] for k in tb]), lambda x, t: 2 + t * x[0]))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.0)
    och.update(0.0)
    K, A = fb.allocate_matrix(dh), fb.allocate_matrix(dh)
    f, gv = ctx.zeros(dh.ndofs), ctx.zeros(dh.ndofs)
    fb.assemble_(fb.start_assemble(K, f), elem, cv)
    fb.assemble_(fb.start_assemble(A, gv), elem, cv)
    data = fb.get_rhs_data(ch, A)
    oA, og_ = O.allocate_matrix(odh), np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oA, og_, "heat", op)
    odata = O.get_rhs_data(och, oA)
    assert abs(data.m - odata[0]) <= 1e-13 * odata[0] and data.nprescribed == len(och.prescribed_dofs)
    assert data.nstored == sum(len(r) for r, _ in odata[1])
    fb.apply_(K, f, ch, applyzero=applyzero)
    fb.apply_(A, None, ch)                      # apply!(A, ch): the matrix once
    fb.apply_rhs_(data, gv, ch, applyzero=applyzero)
    assert close(gv.cpu().numpy(), f.cpu().numpy())[0]
    assert close(gv.cpu().numpy(), O.apply_rhs(odata, og_.copy(), och, applyzero=applyzero))[0]
    assert close(K.nzval.cpu().numpy(), A.nzval.cpu().numpy())[0]        # two atomic assemblies: equal up to summation order
    # next time step: new inhomogeneities on a fresh right-hand side, K untouched
    fb.update_(ch, 1.5)
    och.update(1.5)
    g2 = ctx.zeros(dh.ndofs)
    fb.assemble_(fb.start_assemble(fb.allocate_matrix(dh), g2), elem, cv)
    fb.apply_rhs_(data, g2, ch, applyzero=applyzero)
    assert close(g2.cpu().numpy(), O.apply_rhs(odata, og_.copy(), och, applyzero=applyzero))[0]
    u = ctx.zeros(dh.ndofs)
    fb.cg_(u, A, g2, reltol=1e-13, jacobi=True)
    expect = np.zeros(len(och.prescribed_dofs)) if applyzero else och.inhomogeneities
    assert np.allclose(u.cpu().numpy()[och.prescribed_dofs - 1], expect, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("ct,nel,order,vdim,qo", [
    (fb.Hexahedron, (4, 3, 3), 1, 1, 2), (fb.Hexahedron, (3, 2, 2), 2, 3, 3), (fb.Tetrahedron, (3, 2, 2), 2, 3, 4),
    (fb.Quadrilateral, (5, 4), 2, 1, 3), (fb.Triangle, (5, 4), 1, 2, 2),
])
def test_cellvalues_accessors_on_an_affine_field(ctx, ct, nel, order, vdim, qo):
    # test/test_cellvalues.jl:54-98: a field that is affine in x is reproduced exactly by function_value / function_gradient /
    # function_symmetric_gradient / function_divergence at spatial_coordinate(cv, q, x); volumes = sum detJdV
    # (test/test_utils.jl:130-222); shape_value / shape_gradient / getdetJdV against the oracle's reinit!
    import torch
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    dim = len(nel)
    rv = fb.reinit_batch(cv, g)
    dNdx, dO = O.reinit(ocv, og.nodes[og.cells - 1])
    assert rv.getnquadpoints() == ocv.w.size and rv.getnbasefunctions() == odh.ndofs_per_cell
    rng = np.random.default_rng(4)
    V, G = rng.random(vdim), rng.random((vdim, dim))
    # dof coordinates (interpolation of the multilinear geometry at the reference coordinates of the basis)
    oip, geo = odh.field_ips[0].base, O.Lagrange(og.shape, 1)
    xe = np.stack([np.einsum("j,cjd->cd", geo.value_and_gradient(oip.refcoords[a])[0], og.nodes[og.cells - 1])
                   for a in range(oip.nbase)], axis=1)                        # (nc, nbase, dim)
    ue = (np.einsum("vd,cad->cav", G, xe) + V).reshape(og.ncells, -1)
    ue_t = torch.from_numpy(ue).to(f"cuda:{ctx.device}")
    vol = 0.0
    for q in range(1, rv.getnquadpoints() + 1):
        xq = rv.spatial_coordinate(q).cpu().numpy()
        assert close(xq, np.einsum("j,cjd->cd", ocv.M[q - 1], og.nodes[og.cells - 1]), 1e-14)[0]
        assert close(rv.getdetJdV(q).cpu().numpy(), dO[:, q - 1], 1e-13)[0]
        vol += float(rv.getdetJdV(q).sum())
        val, grad = rv.function_value(q, ue_t).cpu().numpy(), rv.function_gradient(q, ue_t).cpu().numpy()
        exact = np.einsum("vd,cd->cv", G, xq) + V
        if vdim == 1:
            assert np.allclose(val, exact[:, 0], rtol=1e-12, atol=1e-13) and np.allclose(grad, G[0], rtol=0, atol=1e-11)
            assert np.allclose(rv.function_divergence(q, ue_t).cpu().numpy(), G[0].sum(), rtol=0, atol=1e-11)
        else:
            assert np.allclose(val, exact, rtol=1e-12, atol=1e-13) and np.allclose(grad, G, rtol=0, atol=1e-11)
            assert np.allclose(rv.function_symmetric_gradient(q, ue_t).cpu().numpy(), 0.5 * (G + G.T), rtol=0, atol=1e-11)
            assert np.allclose(rv.function_divergence(q, ue_t).cpu().numpy(), np.trace(G), rtol=0, atol=1e-11)
        for i in (1, rv.getnbasefunctions()):
            a, c = (i - 1) // vdim, (i - 1) % vdim
            sg = rv.shape_gradient(q, i).cpu().numpy()
            sv = rv.shape_value(q, i)
            if vdim == 1:
                assert sv == ocv.N[q - 1, a] and close(sg, dNdx[:, q - 1, a, :], 1e-13)[0]
            else:
                assert float(sv[c]) == ocv.N[q - 1, a] and float(sv.sum()) == ocv.N[q - 1, a]
                assert close(sg[:, c, :], dNdx[:, q - 1, a, :], 1e-13)[0] and np.count_nonzero(np.delete(sg, c, axis=1)) == 0
                assert close(rv.shape_divergence(q, i).cpu().numpy(), dNdx[:, q - 1, a, c], 1e-13)[0]
    assert abs(vol - 2.0 ** dim) < 1e-12 * 2.0 ** dim

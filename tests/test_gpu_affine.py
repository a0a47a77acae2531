"""Affine and periodic constraints on the device: AffineConstraint add!/close!, the condensed pattern of
allocate_matrix(dh, ch), apply!(K, f, ch) with `_condense!`, apply!(u, ch), PeriodicDirichlet -- against oracle/affine.py
(pinned on test/test_constraints.jl:323-412) and the reference's `ch_p` goldens (test/test_constraints.jl:1425-1610)."""
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


def line_problem(ncells=10):
    g = fb.generate_grid(fb.Line, (ncells,))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefLine, 1)))
    og = O.generate_grid("line", (ncells,))
    odh = O.DofHandler(og).add("u", O.Lagrange("line", 1)).close()
    return g, dh, og, odh


def test_nonsymmetric_condensation_on_a_dense_pattern(ctx):
    # test/test_constraints.jl:330-341: K = reshape(1:n^2, n, n), two affine constraints, apply!(K, ch) == C' K C on the free dofs
    g, dh, og, odh = line_problem()
    n = dh.ndofs
    ch, och = fb.ConstraintHandler(dh), O.AffineConstraintHandler(odh)
    for c in (fb.AffineConstraint(1, [(3, 2.0)], 0.0), fb.AffineConstraint(2, [(4, 3.0)], 0.0)):
        fb.add_(ch, c)
        och.add(O.AffineConstraint(c.constrained_dof, c.entries, c.b))
    fb.close_(ch)
    och.close()
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs)
    assert ch.dofcoefficients == [list(map(tuple, e)) if e else None for e in och.dofcoefficients]
    oK = O.dense_pattern(n)
    oK.nzval[:] = np.arange(1.0, n * n + 1)
    K = fb.allocate_matrix(dh, oK.colptr, oK.rowval)
    K.nzval.copy_(__import__("torch").from_numpy(oK.nzval))
    fb.apply_(K, None, ch)
    och.apply(oK)
    assert close(K.nzval.cpu().numpy(), oK.nzval, 1e-14)


@pytest.mark.parametrize("acs", [
    [(4, [(7, 1.0)], 0.0)],
    [(2, [(5, 1.0), (6, 2.0)], 1.0)],
    [(2, [(9, 1.0)], 0.0), (3, [(9, 1.0)], 0.0)],
    [(2, [(7, 3.0), (8, 1.0)], -1.0), (4, [(9, -1.0)], 2.0)],
])
@pytest.mark.parametrize("applyzero", [False, True])
def test_affine_constraints_on_a_line(ctx, acs, applyzero):
    # test/test_constraints.jl:343-412: Dirichlet on the left + affine constraints; pattern bit-exact, K / f after apply! entry-wise,
    # and the solution satisfies the constraints
    import torch
    g, dh, og, odh = line_problem()
    n = dh.ndofs
    ch, och = fb.ConstraintHandler(dh), O.AffineConstraintHandler(odh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 0.0))
    och.add(O.Dirichlet("u", og.facetsets["left"], lambda x, t: 0.0))
    for d, e, b in acs:
        fb.add_(ch, fb.AffineConstraint(d, e, b))
        och.add(O.AffineConstraint(d, e, b))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.0)
    och.update(0.0)
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs) and np.array_equal(ch.inhomogeneities, och.inhomogeneities)
    K, oK = fb.allocate_matrix(dh, ch), O.allocate_matrix_condensed(odh, och)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    ke = 2.0 * np.array([[1.0, -1.0], [-1.0, 1.0]])
    fb.scatter_(fb.start_assemble(K, None), np.stack([ke] * g.ncells))
    for c in range(og.ncells):
        O.assemble_cell(oK, None, odh.cell_dofs[c], ke)
    fh = np.zeros(n)
    fh[-1] = 1.0
    f, of = torch.from_numpy(fh).to(K.nzval.device), fh.copy()
    m = fb.apply_(K, f, ch, applyzero=applyzero)
    om = och.apply(oK, of, applyzero=applyzero)
    assert abs(m - om) <= 1e-14 * abs(om)
    assert close(K.nzval.cpu().numpy(), oK.nzval, 1e-13) and close(f.cpu().numpy(), of, 1e-13)
    u = torch.from_numpy(np.linalg.solve(oK.toscipy().toarray(), of)).to(K.nzval.device)
    fb.apply_(u, ch, applyzero=applyzero)
    a = u.cpu().numpy()
    oa = och.apply_vec(np.linalg.solve(oK.toscipy().toarray(), of), applyzero=applyzero)
    assert close(a, oa, 1e-13)
    if not applyzero:
        for d, e, b in acs:
            assert abs(a[d - 1] - (b + sum(v * a[k - 1] for k, v in e))) < 1e-12


def test_nested_affine_constraints_are_rejected(ctx):
    g, dh, og, odh = line_problem()
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.AffineConstraint(1, [(2, 1.0)], 0.0))
    fb.add_(ch, fb.AffineConstraint(2, [(3, 1.0)], 0.0))
    with pytest.raises(fb.FB2Error, match="nested affine constraints currently not supported"):
        fb.close_(ch)


@pytest.mark.parametrize("ct,shape,nel,order", [(fb.Quadrilateral, "quadrilateral", (7, 6), 1), (fb.Triangle, "triangle", (5, 4), 2),
                                                 (fb.Hexahedron, "hexahedron", (4, 3, 3), 1)])
def test_random_affine_constraints_pattern_and_apply(ctx, ct, shape, nel, order):
    """random affine constraints (0..3 masters each) on 2-D / 3-D grids: condensed pattern bit-exact (built on the device from
    the pseudo-cell table), heat K / f after apply! entry-wise against the oracle's `_condense!`"""
    rng = np.random.default_rng(11)
    g = fb.generate_grid(ct, nel).perturb(0.15)
    og = O.perturb_grid(O.generate_grid(shape, nel), nel, (-1.0,) * len(nel), (1.0,) * len(nel), 0.15)
    ip, oip = fb.Lagrange(ct, order), O.Lagrange(shape, order)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    odh = O.DofHandler(og).add("u", oip).close()
    n = dh.ndofs
    slaves = rng.choice(np.arange(1, n + 1), size=max(3, n // 6), replace=False)
    pool = np.setdiff1d(np.arange(1, n + 1), slaves)
    ch, och = fb.ConstraintHandler(dh), O.AffineConstraintHandler(odh)
    for s in slaves:
        masters = rng.choice(pool, size=rng.integers(0, 4), replace=False)
        e = [(int(m), float(rng.random() + 0.5)) for m in masters]
        b = float(rng.random())
        fb.add_(ch, fb.AffineConstraint(int(s), e, b))
        och.add(O.AffineConstraint(int(s), e, b))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.0)
    och.update(0.0)
    K, oK = fb.allocate_matrix(dh, ch), O.allocate_matrix_condensed(odh, och)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    qo = 2 if order == 1 else 3
    cv, ocv = fb.CellValues(fb.QuadratureRule(ct, qo), ip), O.CellValues(O.QuadratureRule(shape, qo), oip)
    f, of = ctx.zeros(n), np.zeros(n)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(k=1.5, source=0.7), cv)
    O.assemble_global(odh, ocv, oK, of, "heat", dict(k=1.5, source=0.7))
    fb.apply_(K, f, ch)
    och.apply(oK, of)
    assert close(K.nzval.cpu().numpy(), oK.nzval) and close(f.cpu().numpy(), of)
    # a matrix without the condensed entries must be refused, not silently mangled
    K0 = fb.allocate_matrix(dh)
    fb.assemble_(fb.start_assemble(K0, f), fb.HeatElement(), cv)
    if K0.nnz < K.nnz:
        with pytest.raises(fb.MissingPatternEntry):
            fb.apply_(K0, f, ch)


@pytest.mark.parametrize("applyzero,golden", [(False, 3.7828270430540893), (True, 0.02672553850330505)])
def test_periodic_dirichlet_reference_golden(ctx, applyzero, golden):
    """test/test_constraints.jl:1425-1610, the ch_p arm: heat on 5 x 5 quadrilaterals with conductivity k = cellid and source
    1 / cellid, bottom tied to top periodically, Dirichlet 0 / 1 on left / right; norm(u_p) is a literal of the reference."""
    import torch
    g = fb.generate_grid(fb.Quadrilateral, (5, 5))
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.PeriodicDirichlet("u", fb.collect_periodic_facets(g, "bottom", "top")))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 0.0))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), lambda x, t: 1.0))
    fb.close_(ch)
    fb.update_(ch, 0.0)
    K = fb.allocate_matrix(dh, ch)
    f = ctx.zeros(dh.ndofs)
    # element!(ke, fe, cv, k = cellid, b = 1 / cellid): element matrices of the unit problem on the device, scaled per cell
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip)
    ea = fb.ElementAssembly(dh, cv)
    Kes, fes = ea.assemble(fb.HeatElement(1.0, 1.0))
    ids = torch.arange(1, g.ncells + 1, dtype=torch.float64, device=Kes.device)
    Kes = Kes * ids.view(-1, 1, 1)
    fes = fes / ids.view(-1, 1)
    fb.scatter_device_(fb.start_assemble(K, f), Kes.contiguous(), fes.contiguous())
    fb.apply_(K, f, ch, applyzero=applyzero)
    u = torch.from_numpy(spla.spsolve(K.tocsc(), f.cpu().numpy())).to(f.device)
    fb.apply_(u, ch, applyzero=applyzero)
    assert abs(float(u.norm()) - golden) <= 1e-10 * golden

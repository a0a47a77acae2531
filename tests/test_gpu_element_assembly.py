"""GPU parity tests of the element-assembly path: element matrices kept on the device, the matrix-free
operator, apply_local! / apply_assemble!.  Oracle: oracle/assemble.py (element_matrices, ea_mul), oracle/constraints.py
(apply_local, apply_assemble), pinned on the reference's "local application of bc" goldens (tests/test_oracle_goldens.py).
Tolerance: 1e-12 relative as for the assembled matrices (test_gpu_parity.py)."""
import numpy as np
import pytest

import ferrite_b200 as fb
import oracle as O
from test_gpu_parity import build, close, displacement, make_element

pytestmark = pytest.mark.gpu

CASES = [
    ("hex-q1-heat", fb.Hexahedron, (5, 4, 3), 1, 1, 2, "heat", {"k": 2.0, "source": 3.0}),
    ("hex-q1-heat-large", fb.Hexahedron, (14, 12, 11), 1, 1, 2, "heat", {}),
    ("quad-q2-mass", fb.Quadrilateral, (5, 5), 2, 1, 3, "mass", {"rho": 1.5}),
    ("tet-p2-heat", fb.Tetrahedron, (3, 3, 2), 2, 1, 2, "heat", {}),
    ("hex-q1-elast", fb.Hexahedron, (5, 4, 3), 1, 3, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    ("hex-q2-elast", fb.Hexahedron, (3, 2, 2), 2, 3, 3, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    ("tri-p2-elast", fb.Triangle, (6, 5), 2, 2, 2, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.5, -1.0)}),
    ("tet-p2-neohooke", fb.Tetrahedron, (3, 2, 2), 2, 3, 4, "neohooke", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
]


@pytest.fixture(scope="module")
def ctx():
    return fb.default_context(0)


def problem(ctx, case):
    import torch
    _, ct, nel, order, vdim, qo, kind, p = case
    left = (0.0,) * len(nel) if kind == "neohooke" else None
    right = (1.0,) * len(nel) if kind == "neohooke" else None
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo, True, left, right)
    elem, op = make_element(kind, p)
    u = ou = None
    if kind == "neohooke":
        ou = displacement(og, odh, vdim)
        u = torch.from_numpy(ou).to(f"cuda:{ctx.device}")
    return g, og, dh, odh, cv, ocv, elem, op, kind, u, ou


def dirichlet(g, og, dh, odh, vdim):
    def val(x, t):
        return [0.01 * x[1] + 0.02 * k + t for k in range(vdim)] if vdim > 1 else 0.3 * x[0] - x[1] + t
    ch, och = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: [0.0] * vdim if vdim > 1 else 0.0))
    och.add(O.Dirichlet("u", og.facetsets["left"], lambda x, t: [0.0] * vdim if vdim > 1 else 0.0))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), val))
    och.add(O.Dirichlet("u", og.facetsets["right"], val))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.5)
    och.update(0.5)
    return ch, och


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_element_matrices_and_operator_match_oracle(ctx, case):
    import torch
    g, og, dh, odh, cv, ocv, elem, op, kind, u, ou = problem(ctx, case)
    ea = fb.ElementAssembly(dh, cv)
    assert (ea.ncells, ea.n) == (og.ncells, odh.ndofs_per_cell)
    Kes, fes = ea.assemble(elem, u=u)
    oKes, ofes = O.element_matrices(odh, ocv, kind, op, u=ou)
    ok, nrm = close(Kes.cpu().numpy().transpose(0, 2, 1), oKes)
    assert ok, nrm
    if kind != "mass":
        ok, nrm = close(fes.cpu().numpy(), ofes)
        assert ok, nrm
    # the operator: y = sum_e P' Ke P x equals the oracle's cell loop and K * x of the oracle's assembled matrix
    x = np.cos(0.37 * np.arange(odh.ndofs)) + 0.1
    y = ea.mul(Kes, torch.from_numpy(x).to(Kes.device))
    oy = O.ea_mul(odh, oKes, x)
    ok, nrm = close(y.cpu().numpy(), oy)
    assert ok, nrm
    oK = O.allocate_matrix(odh)
    O.assemble_global(odh, ocv, oK, np.zeros(odh.ndofs), kind, op, u=ou)
    ok, nrm = close(y.cpu().numpy(), oK.toscipy() @ x, rtol=1e-11)
    assert ok, nrm
    # scattering the stored element matrices gives the assembled matrix (assemble! from device-resident Ke)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    fb.scatter_device_(fb.start_assemble(K, f), Kes, fes)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[4], CASES[5], CASES[7]], ids=lambda c: c[0])
@pytest.mark.parametrize("applyzero", [False, True])
def test_apply_local_and_apply_assemble_match_oracle(ctx, case, applyzero):
    g, og, dh, odh, cv, ocv, elem, op, kind, u, ou = problem(ctx, case)
    vdim = case[4]
    ch, och = dirichlet(g, og, dh, odh, vdim)
    ea = fb.ElementAssembly(dh, cv)
    Kes, fes = ea.assemble(elem, u=u)
    ea.apply_local_(Kes, fes, ch, applyzero=applyzero)
    oKes, ofes = O.element_matrices(odh, ocv, kind, op, u=ou)
    oK, of = O.allocate_matrix(odh), np.zeros(odh.ndofs)
    for c in range(og.ncells):
        O.apply_local(oKes[c], ofes[c], odh.cell_dofs[c], och, applyzero=applyzero)
        O.assemble_cell(oK, of, odh.cell_dofs[c], oKes[c], ofes[c])
    ok, nrm = close(Kes.cpu().numpy().transpose(0, 2, 1), oKes)
    assert ok, nrm
    ok, nrm = close(fes.cpu().numpy(), ofes)
    assert ok, nrm
    # apply_assemble! for the whole cell loop
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    K.nzval.fill_(5.0)
    f.fill_(-3.0)
    fb.apply_assemble_(fb.start_assemble(K, f), ch, elem, cv, u=u, applyzero=applyzero, ea=ea)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, nrm
    # same free block as apply!(K, f, ch); prescribed rows/columns are diagonal (test/test_constraints.jl:1566-1590)
    Ks = fb.allocate_matrix(dh)
    fs = ctx.zeros(dh.ndofs)
    fb.assemble_(fb.start_assemble(Ks, fs), elem, cv, u=u)
    fb.apply_(Ks, fs, ch, applyzero=applyzero)
    A, As = K.tocsc().tocsr(), Ks.tocsc().tocsr()
    pd = och.prescribed_dofs - 1
    fd = np.setdiff1d(np.arange(odh.ndofs), pd)
    ok, nrm = close(A[fd][:, fd].toarray(), As[fd][:, fd].toarray(), rtol=1e-11)
    assert ok, nrm
    ok, nrm = close(f.cpu().numpy()[fd], fs.cpu().numpy()[fd], rtol=1e-10)
    assert ok, nrm
    P = A[pd][:, pd].toarray()
    assert np.array_equal(P, np.diag(np.diag(P)))
    assert abs(A[pd][:, fd]).max() == 0 and abs(A[fd][:, pd]).max() == 0
    expect = np.zeros(len(pd)) if applyzero else och.inhomogeneities
    assert np.allclose(f.cpu().numpy()[pd] / np.diag(P), expect, rtol=1e-13, atol=0)


def test_local_application_of_bc_golden_through_the_gpu_path(ctx):
    # test/test_constraints.jl:1425-1610: conductivity k = cellid, source 1/cellid, norm(u) goldens
    import torch
    g = fb.generate_grid(fb.Quadrilateral, (5, 5))
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip)
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 0))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), lambda x, t: 1))
    fb.close_(ch)
    fb.update_(ch, 0.0)
    ea = fb.ElementAssembly(dh, cv)
    for azero, golden in ((False, 3.8249286998373586), (True, 0.06401424182205259)):
        Kes, fes = ea.assemble(fb.HeatElement(1.0, 1.0))
        ids = torch.arange(1, g.ncells + 1, dtype=torch.float64, device=Kes.device)
        Kes *= ids[:, None, None]
        fes /= ids[:, None]
        ea.apply_local_(Kes, fes, ch, applyzero=azero)
        K = fb.allocate_matrix(dh)
        f = ctx.zeros(dh.ndofs)
        fb.scatter_device_(fb.start_assemble(K, f), Kes, fes)
        u = np.linalg.solve(K.tocsc().toarray(), f.cpu().numpy())
        assert abs(np.linalg.norm(u) - golden) < 1e-12 * golden, (azero, np.linalg.norm(u))


def _oracle_apply_assemble(og, odh, ocv, och, kind, op, applyzero=False):
    oKes, ofes = O.element_matrices(odh, ocv, kind, op)
    oK, of = O.allocate_matrix(odh), np.zeros(odh.ndofs)
    for c in range(og.ncells):
        O.apply_assemble(oK, of, och, odh.cell_dofs[c], oKes[c], ofes[c], applyzero=applyzero)
    return oK, of


def test_apply_assemble_split_cache_follows_update_and_coordinates(ctx):
    # the boundary-layer split is cached per ConstraintHandler: update!(ch, t) and new node coordinates must show up in
    # the next call; a ConstraintHandler with other prescribed dofs rebuilds the split
    ct, nel = fb.Hexahedron, (9, 8, 7)
    g, og, dh, odh, cv, ocv = build(ct, nel, 1, 1, 2)
    elem, op = make_element("heat", {"k": 1.5, "source": 0.5})
    ch, och = dirichlet(g, og, dh, odh, 1)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    a = fb.start_assemble(K, f)
    ea = fb.apply_assemble_(a, ch, elem, cv)
    oK, of = _oracle_apply_assemble(og, odh, ocv, och, "heat", op)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    # new time, new coordinates
    fb.update_(ch, 2.0)
    och.update(2.0)
    og.nodes[:] = og.nodes * np.array([1.0, 1.1, 0.9]) + 0.05
    g.set_coordinates(og.nodes)
    fb.apply_assemble_(fb.start_assemble(K, f), ch, elem, cv, ea=ea)
    oK, of = _oracle_apply_assemble(og, odh, ocv, och, "heat", op)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    # another ConstraintHandler through the same ElementAssembly
    ch2, och2 = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    fb.add_(ch2, fb.Dirichlet("u", fb.getfacetset(g, "top"), lambda x, t: 2.0 + x[0]))
    och2.add(O.Dirichlet("u", og.facetsets["top"], lambda x, t: 2.0 + x[0]))
    fb.close_(ch2)
    och2.close()
    fb.apply_assemble_(fb.start_assemble(K, f), ch2, elem, cv, ea=ea, applyzero=True)
    oK, of = _oracle_apply_assemble(og, odh, ocv, och2, "heat", op, applyzero=True)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    # fillzero = false accumulates
    a2 = fb.start_assemble(K, f, fillzero=False)
    fb.apply_assemble_(a2, ch2, elem, cv, ea=ea, applyzero=True)
    assert close(K.nzval.cpu().numpy(), 2 * oK.nzval)[0] and close(f.cpu().numpy(), 2 * of)[0]


def test_apply_assemble_edge_cases(ctx):
    # every cell on the boundary (no interior launch) and no prescribed dof at all (plain assembly)
    g, og, dh, odh, cv, ocv = build(fb.Quadrilateral, (2, 3), 1, 1, 2)
    elem, op = make_element("heat", {})
    ch, och = dirichlet(g, og, dh, odh, 1)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    fb.apply_assemble_(fb.start_assemble(K, f), ch, elem, cv)
    oK, of = _oracle_apply_assemble(og, odh, ocv, och, "heat", op)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    empty = fb.ConstraintHandler(dh)
    fb.close_(empty)
    fb.apply_assemble_(fb.start_assemble(K, f), empty, elem, cv)
    oK, of = O.allocate_matrix(odh), np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, "heat", op)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]

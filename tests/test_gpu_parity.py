"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): dof numbering, colptr and rowval bit-exact; nzval / f within 1e-12
relative in FP64 (atomic summation order).  Entry-wise relative error is ill-defined where K is
analytically zero (e.g. edge-adjacent entries of the trilinear Laplacian), so the entry-wise check is
|a-b| <= RTOL*max(|a|,|b|) + RTOL*max|a| and the norm-wise error is checked as well.
"""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import ferrite_b200 as fb
import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12
SHAPE = {fb.Triangle: "triangle", fb.Quadrilateral: "quadrilateral", fb.Tetrahedron: "tetrahedron",
         fb.Hexahedron: "hexahedron", fb.Line: "line"}


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max() if b.size else 0.0
    ok = np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * scale)
    nrm = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    return bool(ok and nrm <= rtol), nrm


@pytest.fixture(scope="module")
def ctx():
    return fb.default_context(0)


def build(ct, nel, order, vdim, qr_order, perturb=True, left=None, right=None):
    dim = len(nel)
    left = left if left is not None else (-1.0,) * dim
    right = right if right is not None else (1.0,) * dim
    g = fb.generate_grid(ct, nel, left, right)
    og = O.generate_grid(SHAPE[ct], nel, left, right)
    if perturb:
        g.perturb(0.2)
        O.perturb_grid(og, nel, left, right, 0.2)
    ip = fb.Lagrange(ct, order) ** vdim
    oip = O.Lagrange(SHAPE[ct], order)
    oip = oip ** vdim if vdim > 1 else oip
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    odh = O.DofHandler(og).add("u", oip).close()
    cv = fb.CellValues(fb.QuadratureRule(ct, qr_order), ip)
    ocv = O.CellValues(O.QuadratureRule(SHAPE[ct], qr_order), oip)
    return g, og, dh, odh, cv, ocv


def test_pattern_bit_exact_many(ctx):
    cases = [
        (fb.Quadrilateral, (7, 5), 1, 1), (fb.Quadrilateral, (5, 4), 2, 2), (fb.Triangle, (6, 5), 2, 1),
        (fb.Hexahedron, (5, 4, 3), 1, 1), (fb.Hexahedron, (4, 3, 3), 1, 3), (fb.Hexahedron, (3, 3, 2), 2, 3),
        (fb.Tetrahedron, (3, 3, 2), 1, 3), (fb.Tetrahedron, (3, 2, 2), 2, 3), (fb.Line, (9,), 2, 1),
    ]
    for ct, nel, order, vdim in cases:
        g = fb.generate_grid(ct, nel)
        dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, order) ** vdim))
        K = fb.allocate_matrix(dh)
        og = O.generate_grid(SHAPE[ct], nel)
        oip = O.Lagrange(SHAPE[ct], order)
        odh = O.DofHandler(og).add("u", oip ** vdim if vdim > 1 else oip).close()
        oK = O.allocate_matrix(odh)
        assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
        assert K.nnz == oK.nnz, (ct, nel, order, vdim)
        assert np.array_equal(K.colptr, oK.colptr)
        assert np.array_equal(K.rowval, oK.rowval)


def test_pattern_two_fields_and_from_host(ctx):
    g = fb.generate_grid(fb.Hexahedron, (3, 2, 2))
    h1, h2 = fb.Lagrange(fb.RefHexahedron, 1), fb.Lagrange(fb.RefHexahedron, 2)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", h2 ** 3), "p", h1))
    K = fb.allocate_matrix(dh)
    og = O.generate_grid("hexahedron", (3, 2, 2))
    odh = O.DofHandler(og).add("u", O.Lagrange("hexahedron", 2) ** 3).add("p", O.Lagrange("hexahedron", 1)).close()
    oK = O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    # arrays-in mode round trip
    K2 = fb.allocate_matrix(dh, oK.colptr, oK.rowval)
    assert np.array_equal(K2.colptr, oK.colptr) and np.array_equal(K2.rowval, oK.rowval)
    with pytest.raises(fb.FB2Error):
        fb.allocate_matrix(dh, oK.colptr, oK.rowval[::-1].copy())


ASSEMBLY_CASES = [
    # (celltype, nel, order, vdim, qr_order, element, oracle params, fb element)
    ("quad-q1-heat", fb.Quadrilateral, (13, 9), 1, 1, 2, "heat", {"k": 1.3, "source": 0.7}),
    ("quad-q2-heat", fb.Quadrilateral, (7, 6), 2, 1, 3, "heat", {"k": 1.0, "source": 1.0}),
    ("tri-p1-heat", fb.Triangle, (9, 8), 1, 1, 1, "heat", {}),
    ("tri-p2-heat", fb.Triangle, (8, 7), 2, 1, 2, "heat", {}),
    ("hex-q1-heat", fb.Hexahedron, (9, 8, 7), 1, 1, 2, "heat", {"k": 2.0, "source": 3.0}),
    ("hex-q2-heat", fb.Hexahedron, (4, 3, 3), 2, 1, 3, "heat", {}),
    ("tet-p1-heat", fb.Tetrahedron, (4, 4, 3), 1, 1, 2, "heat", {}),
    ("tet-p2-heat", fb.Tetrahedron, (3, 3, 3), 2, 1, 2, "heat", {}),
    ("tet-p2-heat-q3", fb.Tetrahedron, (3, 2, 3), 2, 1, 3, "heat", {}),
    # sizes above the tile-kernel threshold (k_tile_scalar: Morton tiles, shared-memory aggregation)
    ("hex-q1-heat-tiles", fb.Hexahedron, (17, 13, 11), 1, 1, 2, "heat", {"k": 2.0, "source": 3.0}),
    ("quad-q1-heat-tiles", fb.Quadrilateral, (53, 41), 1, 1, 2, "heat", {"k": 1.3, "source": 0.7}),
    ("tri-p2-heat-tiles", fb.Triangle, (31, 27), 2, 1, 2, "heat", {}),
    ("tet-p2-heat-tiles", fb.Tetrahedron, (7, 6, 5), 2, 1, 2, "heat", {}),
    ("tet-p1-heat-tiles", fb.Tetrahedron, (8, 7, 6), 1, 1, 2, "heat", {}),
    ("hex-q1-mass-tiles", fb.Hexahedron, (13, 12, 11), 1, 1, 2, "mass", {"rho": 2.5}),
    ("hex-q1-mass", fb.Hexahedron, (5, 4, 4), 1, 1, 2, "mass", {"rho": 2.5}),
    ("quad-q2-mass", fb.Quadrilateral, (5, 5), 2, 1, 3, "mass", {"rho": 1.0}),
    ("quad-q1-elast", fb.Quadrilateral, (8, 7), 1, 2, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, -1.0)}),
    ("tri-p2-elast", fb.Triangle, (6, 5), 2, 2, 2, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.5, -1.0)}),
    ("hex-q1-elast", fb.Hexahedron, (6, 5, 4), 1, 3, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    ("hex-q2-elast", fb.Hexahedron, (3, 3, 2), 2, 3, 3, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    ("tet-p1-elast", fb.Tetrahedron, (3, 3, 3), 1, 3, 1, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
    ("tet-p2-elast", fb.Tetrahedron, (3, 2, 2), 2, 3, 4, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
    ("tet-p1-neohooke", fb.Tetrahedron, (3, 3, 3), 1, 3, 1, "neohooke", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
    ("tet-p2-neohooke", fb.Tetrahedron, (3, 2, 2), 2, 3, 4, "neohooke", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
    ("hex-q1-neohooke", fb.Hexahedron, (4, 3, 3), 1, 3, 2, "neohooke", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
]


def make_element(kind, p):
    if kind == "heat":
        return fb.HeatElement(**p), dict(k=p.get("k", 1.0), source=p.get("source", 1.0))
    if kind == "mass":
        return fb.MassElement(**p), dict(p)
    lam, mu = O.lame(p["E"], p["nu"])
    cls = fb.ElasticityElement if kind == "elasticity" else fb.NeoHookeElement
    return cls(lam=lam, mu=mu, b=p["b"]), {"lambda": lam, "mu": mu, "b": p["b"]}


def displacement(og, odh, vdim):
    """Deterministic smooth state u = 0.05 sin(pi x)-type field evaluated at the dofs (config 4)."""
    u = np.zeros(odh.ndofs)
    base = odh.field_ips[0].base
    geo = O.Lagrange(og.shape, 1)
    for ci in range(og.ncells):
        xc = og.nodes[og.cells[ci] - 1]
        for a in range(base.nbase):
            M, _ = geo.value_and_gradient(base.refcoords[a])
            x = M @ xc
            for c in range(vdim):
                u[odh.cell_dofs[ci, a * vdim + c] - 1] = 0.05 * np.sin(np.pi * x[c] + 0.3 * c) * np.cos(0.5 * x[(c + 1) % len(x)])
    return u


@pytest.mark.parametrize("case", ASSEMBLY_CASES, ids=[c[0] for c in ASSEMBLY_CASES])
@pytest.mark.parametrize("scatter", ["atomic", "colored"])
def test_assembly_matches_oracle(ctx, case, scatter):
    _, ct, nel, order, vdim, qo, kind, p = case
    left = (0.0,) * len(nel) if kind == "neohooke" else None
    right = (1.0,) * len(nel) if kind == "neohooke" else None
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo, True, left, right)
    elem, op = make_element(kind, p)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    u = ou = None
    if kind == "neohooke":
        import torch
        ou = displacement(og, odh, vdim)
        u = torch.from_numpy(ou).to(f.device)
    O.assemble_global(odh, ocv, oK, of, kind, op, u=ou)
    variants = [0, 1, 2, 5, 6, 7, 8, 9, 12, 30] if kind in ("heat", "mass") else [0]   # 0 default (marching tiles on generated Q1 hexahedra, else per-cell kernel), 30 per-cell kernel (x face merge), 1 block kernel, 2 unrolled, 5 tile kernel, 6 x+y face merge
    for variant in variants:
        a = fb.start_assemble(K, f, scatter=scatter)
        a.variant = variant
        K.nzval.fill_(123.0)      # start_assemble must zero-fill
        f.fill_(-7.0)
        fb.assemble_(a, elem, cv, u=u)
        fb.finish_assemble(a)
        ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
        assert ok, f"nzval mismatch (variant {variant}): norm-wise {nrm:.3e}"
        if kind != "mass":
            ok, nrm = close(f.cpu().numpy(), of)
            assert ok, f"f mismatch (variant {variant}): norm-wise {nrm:.3e}"


@pytest.mark.parametrize("nel", [(8, 4, 3), (9, 5, 7), (17, 13, 11), (3, 2, 1), (24, 10, 9)])
@pytest.mark.parametrize("lz", ["1", "3", "", "3,1", "4,2"])
def test_marching_tile_kernel(ctx, nel, lz, monkeypatch):
    """k_march_hex (default for generate_grid Q1 hexahedra): full / partial tiles, chunk lengths (FB2_MARCH_LZ), zero fill
    (plain stores for tile-interior columns) and fillzero=false (REDs everywhere), against the oracle and the per-cell kernel."""
    monkeypatch.delenv("FB2_MARCH_LT", raising=False)
    if lz:
        monkeypatch.setenv("FB2_MARCH_LZ", lz.split(",")[0])
        if "," in lz:                                     # short chunks at the end of the launch (tail of the big launches)
            monkeypatch.setenv("FB2_MARCH_LT", lz.split(",")[1])
    else:
        monkeypatch.delenv("FB2_MARCH_LZ", raising=False)
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, nel, 1, 1, 2, True)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, "heat", dict(k=1.7, source=0.3))
    elem = fb.HeatElement(k=1.7, source=0.3)
    a = fb.start_assemble(K, f)
    K.nzval.fill_(55.0)
    f.fill_(-3.0)
    fb.assemble_(a, elem, cv)
    fb.finish_assemble(a)
    assert fb.last_kernel() == "k_march_hex"
    nz, fv = K.nzval.cpu().numpy().copy(), f.cpu().numpy().copy()
    assert close(nz, oK.nzval)[0] and close(fv, of)[0]
    a2 = fb.start_assemble(K, f, fillzero=False)     # accumulate onto the first result
    fb.assemble_(a2, elem, cv)
    fb.finish_assemble(a2)
    assert close(K.nzval.cpu().numpy(), 2 * oK.nzval)[0] and close(f.cpu().numpy(), 2 * of)[0]
    a3 = fb.start_assemble(K, f)
    a3.variant = 30                                   # thread-per-cell kernel
    fb.assemble_(a3, elem, cv)
    fb.finish_assemble(a3)
    assert fb.last_kernel() == "k_cell_scalar"
    assert close(K.nzval.cpu().numpy(), nz)[0] and close(f.cpu().numpy(), fv)[0]


def test_marching_tile_kernel_after_renumbering_and_detj_error(ctx):
    """the accumulator window follows the global column layout, so any dof numbering works; det(J) <= 0 is reported"""
    nel = (10, 9, 6)
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, nel, 1, 1, 2, True)
    rng = np.random.default_rng(5)
    perm = rng.permutation(dh.ndofs) + 1
    fb.renumber_(dh, perm)
    O.renumber(odh, perm)
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, "heat")
    a = fb.start_assemble(K, f)
    fb.assemble_(a, fb.HeatElement(), cv)
    fb.finish_assemble(a)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    xyz = g.nodes.copy()
    xyz[:, 0] *= -1.0                                 # mirrored cells: det(J) < 0 everywhere
    g.set_coordinates(xyz)
    with pytest.raises(fb.DetJNotPositive):
        a = fb.start_assemble(K, f)
        fb.assemble_(a, fb.HeatElement(), cv)
        fb.finish_assemble(a)


@pytest.mark.parametrize("nel", [(4, 4, 2), (9, 6, 5), (13, 7, 11), (3, 2, 1), (17, 10, 9)])
@pytest.mark.parametrize("lz", ["1", "3", ""])
def test_marching_tile_kernel_elasticity(ctx, nel, lz, monkeypatch):
    """k_march_vec (default for isotropic elasticity on generate_grid Q1 hexahedra, BASELINE.json configs[4]'s element): full /
    partial tiles, chunk lengths, zero fill (bulk stores for tile-interior columns) and fillzero=false (reduce-adds everywhere),
    against the oracle and the warp-per-cell kernel."""
    if lz:
        monkeypatch.setenv("FB2_MARCH_LZ", lz)
    else:
        monkeypatch.delenv("FB2_MARCH_LZ", raising=False)
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, nel, 1, 3, 2, True)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    lam, mu = O.lame(10.0, 0.3)
    b = (0.1, 0.2, -1.0)
    O.assemble_global(odh, ocv, oK, of, "elasticity", {"lambda": lam, "mu": mu, "b": b})
    elem = fb.ElasticityElement(lam=lam, mu=mu, b=b)
    a = fb.start_assemble(K, f)
    K.nzval.fill_(55.0)
    f.fill_(-3.0)
    fb.assemble_(a, elem, cv)
    fb.finish_assemble(a)
    assert fb.last_kernel() == "k_march_vec"
    nz, fv = K.nzval.cpu().numpy().copy(), f.cpu().numpy().copy()
    ok, nrm = close(nz, oK.nzval)
    assert ok, f"nzval mismatch: norm-wise {nrm:.3e}"
    ok, nrm = close(fv, of)
    assert ok, f"f mismatch: norm-wise {nrm:.3e}"
    a2 = fb.start_assemble(K, f, fillzero=False)     # accumulate onto the first result
    fb.assemble_(a2, elem, cv)
    fb.finish_assemble(a2)
    assert close(K.nzval.cpu().numpy(), 2 * oK.nzval)[0] and close(f.cpu().numpy(), 2 * of)[0]
    a3 = fb.start_assemble(K, f)
    a3.variant = 32                                   # warp-per-cell kernel
    fb.assemble_(a3, elem, cv)
    fb.finish_assemble(a3)
    assert fb.last_kernel() == "k_cell_syrk"
    assert close(K.nzval.cpu().numpy(), nz)[0] and close(f.cpu().numpy(), fv)[0]
    a4 = fb.start_assemble(K, None)                   # K only
    fb.assemble_(a4, elem, cv)
    fb.finish_assemble(a4)
    assert fb.last_kernel() == "k_march_vec"
    assert close(K.nzval.cpu().numpy(), nz, 1e-13)[0]


def test_marching_tile_kernel_elasticity_after_renumbering_and_detj_error(ctx):
    """the window follows the global column layout (three columns per node, wherever the numbering puts them); det(J) <= 0 is
    reported"""
    nel = (7, 9, 6)
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, nel, 1, 3, 2, True)
    rng = np.random.default_rng(7)
    perm = rng.permutation(dh.ndofs) + 1
    fb.renumber_(dh, perm)
    O.renumber(odh, perm)
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    lam, mu = O.lame(200e9, 0.3)
    O.assemble_global(odh, ocv, oK, of, "elasticity", {"lambda": lam, "mu": mu, "b": (0.0, 0.0, -1.0)})
    elem = fb.ElasticityElement(lam=lam, mu=mu, b=(0.0, 0.0, -1.0))
    a = fb.start_assemble(K, f)
    fb.assemble_(a, elem, cv)
    fb.finish_assemble(a)
    assert fb.last_kernel() == "k_march_vec"
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    # a component-wise numbering: the three columns of a node are far apart in nzval
    fb.renumber_(dh, fb.DofOrder.ComponentWise())
    K2 = fb.allocate_matrix(dh)
    f2 = ctx.zeros(dh.ndofs)
    a = fb.start_assemble(K2, f2)
    fb.assemble_(a, elem, cv)
    fb.finish_assemble(a)
    assert fb.last_kernel() == "k_march_vec"
    nz2, fv2 = K2.nzval.cpu().numpy().copy(), f2.cpu().numpy().copy()
    a = fb.start_assemble(K2, f2)
    a.variant = 32
    fb.assemble_(a, elem, cv)
    fb.finish_assemble(a)
    assert fb.last_kernel() == "k_cell_syrk"
    assert close(nz2, K2.nzval.cpu().numpy())[0] and close(fv2, f2.cpu().numpy())[0]
    xyz = g.nodes.copy()
    xyz[:, 0] *= -1.0                                 # mirrored cells: det(J) < 0 everywhere
    g.set_coordinates(xyz)
    with pytest.raises(fb.DetJNotPositive):
        a = fb.start_assemble(K2, f2)
        fb.assemble_(a, elem, cv)
        fb.finish_assemble(a)


def test_colored_is_bitwise_reproducible_and_coloring_valid(ctx):
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (7, 6, 5), 1, 1, 2)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    a = fb.start_assemble(K, f, scatter="colored")
    ncol, col = a.coloring(cv)
    assert 8 <= ncol <= 16
    # validity: no two cells of one colour share a node (src/Grid/coloring.jl)
    cells = g.cells
    for c in range(ncol):
        nodes = cells[col == c].ravel()
        assert len(np.unique(nodes)) == len(nodes)
    runs = []
    for _ in range(3):
        a = fb.start_assemble(K, f, scatter="colored")
        fb.assemble_(a, fb.HeatElement(), cv)
        fb.finish_assemble(a)
        runs.append((K.nzval.cpu().numpy().copy(), f.cpu().numpy().copy()))
    assert all(np.array_equal(runs[0][0], r[0]) and np.array_equal(runs[0][1], r[1]) for r in runs[1:])


@pytest.mark.parametrize("nel", [(6, 5), (41, 33)])      # per-cell kernel / tile kernel
def test_fillzero_false_accumulates(ctx, nel):
    g, og, dh, odh, cv, ocv = build(fb.Quadrilateral, nel, 1, 1, 2)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    a = fb.start_assemble(K, f)
    a.variant = 5 if nel[0] > 10 else 0
    fb.assemble_(a, fb.HeatElement(), cv)
    once = K.nzval.clone()
    a2 = fb.start_assemble(K, f, fillzero=False)
    a2.variant = a.variant
    fb.assemble_(a2, fb.HeatElement(), cv)
    fb.finish_assemble(a2)
    assert np.allclose(K.nzval.cpu().numpy(), 2 * once.cpu().numpy(), rtol=1e-14)


def test_two_passes_on_one_assembler_accumulate(ctx):
    """start_assemble zeroes K and f once; assemble! calls on the returned assembler add up (src/assembler.jl:287-291,
    322-331): a = start_assemble(K, f); heat pass; mass pass  ==  K_heat + K_mass."""
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (9, 6, 5), 1, 1, 2)
    K, f = fb.allocate_matrix(dh), ctx.zeros(dh.ndofs)
    oK1, oK2, of1 = O.allocate_matrix(odh), O.allocate_matrix(odh), np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK1, of1, "heat", dict(k=2.0, source=3.0))
    O.assemble_global(odh, ocv, oK2, None, "mass", dict(rho=0.5))
    K.nzval.fill_(9.0)
    f.fill_(9.0)
    a = fb.start_assemble(K, f)
    fb.assemble_(a, fb.HeatElement(k=2.0, source=3.0), cv)
    fb.assemble_(a, fb.MassElement(rho=0.5), cv)
    fb.finish_assemble(a)
    assert close(K.nzval.cpu().numpy(), oK1.nzval + oK2.nzval)[0] and close(f.cpu().numpy(), of1)[0]


def test_host_buffer_entry_point(ctx):
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (6, 5, 4), 1, 1, 2)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, "heat")
    nz = np.full(K.nnz, 9.0)
    fh = np.full(dh.ndofs, 9.0)
    a = fb.start_assemble(K, None)
    fb.assemble_host(a, fb.HeatElement(), cv, nz, fh)
    assert close(nz, oK.nzval)[0] and close(fh, of)[0]


def test_arrays_in_mode_matches_native(ctx):
    """Arrays-in entry mode: grid, cell_dofs, pattern, tables and constraints supplied by the caller
    (here: by the oracle standing in for Ferrite.jl) give the same matrix as the native mode."""
    nel = (5, 4, 3)
    og = O.perturb_grid(O.generate_grid("hexahedron", nel), nel, (-1,) * 3, (1,) * 3, 0.2)
    oip = O.Lagrange("hexahedron", 1) ** 3
    odh = O.DofHandler(og).add("u", oip).close()
    oK = O.allocate_matrix(odh)
    ocv = O.CellValues(O.QuadratureRule("hexahedron", 2), oip)
    g = fb.Grid.from_arrays(fb.Hexahedron, og.cells, og.nodes)
    ip = fb.Lagrange(fb.RefHexahedron, 1) ** 3
    dh = fb.DofHandler.from_arrays(g, [("u", ip)], odh.ndofs, odh.cell_dofs)
    K = fb.allocate_matrix(dh, oK.colptr, oK.rowval)
    cv = fb.CellValues(None, ip, tables=(ocv.N, ocv.dNdxi, ocv.M, ocv.dMdxi, ocv.w))
    f = ctx.zeros(dh.ndofs)
    lam, mu = O.lame(200e9, 0.3)
    a = fb.start_assemble(K, f)
    fb.assemble_(a, fb.ElasticityElement(lam=lam, mu=mu, b=(0, 0, -1.0)), cv)
    fb.finish_assemble(a)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, "elasticity", {"lambda": lam, "mu": mu, "b": (0, 0, -1.0)})
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0]
    assert close(f.cpu().numpy(), of)[0]
    # constraints adopted from arrays + apply on a pattern that is not flagged structurally symmetric
    och = O.ConstraintHandler(odh)
    och.add(O.Dirichlet("u", og.facetsets["left"], lambda x, t: (0.0, 0.0, 0.0)))
    och.add(O.Dirichlet("u", og.facetsets["right"], lambda x, t: (0.0, 0.0, 0.01 * x[1]), [1, 2, 3]))
    och.close()
    ch = fb.ConstraintHandler.from_arrays(dh, och.prescribed_dofs, och.inhomogeneities)
    m = fb.apply_(K, f, ch)
    om = och.apply(oK, of)
    assert abs(m - om) <= 1e-13 * abs(om)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0]
    assert close(f.cpu().numpy(), of)[0]


@pytest.mark.parametrize("ct,nel,order,vdim,qo,kind,p", [
    (fb.Quadrilateral, (9, 8), 1, 1, 2, "heat", {}),
    (fb.Hexahedron, (5, 4, 4), 1, 3, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    (fb.Hexahedron, (3, 2, 2), 2, 3, 3, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    (fb.Tetrahedron, (3, 3, 2), 2, 1, 2, "heat", {}),
])
@pytest.mark.parametrize("applyzero", [False, True])
def test_apply_dirichlet_matches_oracle(ctx, ct, nel, order, vdim, qo, kind, p, applyzero):
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    elem, op = make_element(kind, p)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    fb.assemble_(fb.start_assemble(K, f), elem, cv)
    O.assemble_global(odh, ocv, oK, of, kind, op)

    def val(x, t):
        return [0.01 * x[1] + 0.02 * k + t for k in range(vdim)] if vdim > 1 else 0.3 * x[0] - x[1]
    ch, och = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: [0.0] * vdim if vdim > 1 else 0.0))
    och.add(O.Dirichlet("u", og.facetsets["left"], lambda x, t: [0.0] * vdim if vdim > 1 else 0.0))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), val))
    och.add(O.Dirichlet("u", og.facetsets["right"], val))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.5)
    och.update(0.5)
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)
    m = fb.apply_(K, f, ch, applyzero=applyzero)
    om = och.apply(oK, of, applyzero=applyzero)
    assert abs(m - om) <= 1e-13 * abs(om)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, nrm
    # apply!(u, ch)
    u = ctx.zeros(dh.ndofs)
    fb.apply_(u, ch)
    ou = och.apply_vec(np.zeros(odh.ndofs))
    assert np.array_equal(u.cpu().numpy(), ou)


def test_kat_assemble_and_apply_literal(ctx):
    # test/test_assembler_extensions.jl:44-86 through the scatter-only entry point
    g = fb.generate_grid(fb.Line, (2,))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefLine, 1)))
    K = fb.allocate_matrix(dh)
    assert list(K.colptr) == [1, 3, 6, 8] and list(K.rowval) == [1, 2, 1, 2, 3, 2, 3]
    f = ctx.zeros(3)
    ke = np.array([[-1.0, 1.0], [2.0, -1.0]])
    fe = np.array([1.0, 2.0])
    # the reference test scatters cell 2 with dofs [3, 2]; our cell_dofs are [2, 3] -> permute ke/fe accordingly
    P = np.array([1, 0])
    a = fb.start_assemble(K, f)
    fb.scatter_(a, np.stack([ke, ke[np.ix_(P, P)]]), np.stack([fe, fe[P]]))
    fb.finish_assemble(a)
    assert np.allclose(K.tocsc().toarray(), [[-1, 1, 0], [2, -2, 2], [0, 1, -1]], rtol=1e-15)
    assert np.allclose(f.cpu().numpy(), [1.0, 4.0, 1.0], rtol=1e-15)
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 1))
    fb.close_(ch)
    fb.apply_(K, f, ch)
    assert np.allclose(K.tocsc().toarray(), [[4 / 3, 0, 0], [0, -2, 2], [0, 1, -1]], rtol=1e-15)
    assert np.allclose(f.cpu().numpy(), [4 / 3, 2.0, 1.0], rtol=1e-15)


def test_scatter_fake_element_golden(ctx):
    # test/test_assemble.jl:304-357: Ke[i,j] = sin(d_i d_j / 100), fe[i] = cos(d_i) on Q2 10x10 quads
    g = fb.generate_grid(fb.Quadrilateral, (10, 10))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefQuadrilateral, 2)))
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    cd = dh.cell_dofs.astype(np.float64)
    Ke = np.sin(cd[:, :, None] * cd[:, None, :] / 100)
    fe = np.cos(cd)
    a = fb.start_assemble(K, f)
    fb.scatter_(a, Ke, fe)
    fb.finish_assemble(a)
    og = O.generate_grid("quadrilateral", (10, 10))
    odh = O.DofHandler(og).add("u", O.Lagrange("quadrilateral", 2)).close()
    oK = O.allocate_matrix(odh)
    of = np.zeros(odh.ndofs)
    for c in range(og.ncells):
        O.assemble_cell(oK, of, odh.cell_dofs[c], Ke[c], fe[c])
    assert np.allclose(K.nzval.cpu().numpy(), oK.nzval, rtol=1e-14, atol=0)     # the reference's own tolerance
    assert np.allclose(f.cpu().numpy(), of, rtol=1e-14, atol=1e-14)


def test_missing_entry_and_zero_skip(ctx):
    # test/test_assemble.jl:171-214: zeros aimed at missing entries are skipped, non-zeros are an error
    g = fb.generate_grid(fb.Line, (3,))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefLine, 1)))
    # a pattern without the off-diagonal couplings
    n = dh.ndofs
    K = fb.allocate_matrix(dh, np.arange(1, n + 2), np.arange(1, n + 1))
    a = fb.start_assemble(K, None)
    Ke = np.zeros((3, 2, 2))
    Ke[:, 0, 0] = Ke[:, 1, 1] = 1.0
    fb.scatter_(a, Ke)
    assert np.allclose(K.nzval.cpu().numpy(), [1, 2, 2, 1])
    fb.scatter_(a, Ke)                      # a second assemble! on the same assembler adds (start_assemble zeroes once)
    assert np.allclose(K.nzval.cpu().numpy(), [2, 4, 4, 2])
    Ke[1, 0, 1] = 5.0
    with pytest.raises(fb.MissingPatternEntry):
        fb.scatter_(fb.start_assemble(K, None), Ke)


@pytest.mark.parametrize("vdim", [1, 3])
def test_marching_kernels_on_an_incomplete_pattern(ctx, vdim):
    """assemble! on a pattern that lacks entries (src/assembler.jl:459-467) through the marching-tile kernels (their CHECK
    instantiations): zeros aimed at missing entries are skipped, a non-zero is an error"""
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (9, 6, 5), 1, vdim, 2, True)
    K0 = fb.allocate_matrix(dh)
    colptr, rowval = K0.colptr.copy(), K0.rowval.copy()
    # drop the first strictly-lower entry of three interior columns
    drop = []
    for col in (dh.ndofs // 3, dh.ndofs // 2, dh.ndofs - 7):
        lo, hi = colptr[col] - 1, colptr[col + 1] - 1
        below = [p for p in range(lo, hi) if rowval[p] > col + 1]
        drop.append(below[0])
    keep = np.ones(len(rowval), bool)
    keep[drop] = False
    counts = np.diff(colptr).copy()
    for p in drop:
        counts[np.searchsorted(colptr - 1, p, side="right") - 1] -= 1
    colptr2 = np.concatenate(([1], 1 + np.cumsum(counts)))
    K = fb.allocate_matrix(dh, colptr2, rowval[keep])
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    oK = O.allocate_matrix(odh)
    kernel = "k_march_hex" if vdim == 1 else "k_march_vec"
    if vdim == 1:
        zero, full = fb.HeatElement(k=0.0, source=0.7), fb.HeatElement(k=1.0, source=0.7)
        O.assemble_global(odh, ocv, oK, of, "heat", dict(k=0.0, source=0.7))
    else:
        zero, full = fb.ElasticityElement(lam=0.0, mu=0.0, b=(0.1, 0.2, -1.0)), fb.ElasticityElement(lam=1.0, mu=0.5, b=(0.1, 0.2, -1.0))
        O.assemble_global(odh, ocv, oK, of, "elasticity", {"lambda": 0.0, "mu": 0.0, "b": (0.1, 0.2, -1.0)})
    K.nzval.fill_(3.0)
    a = fb.start_assemble(K, f)
    fb.assemble_(a, zero, cv)
    fb.finish_assemble(a)                              # every element matrix is exactly zero: nothing to complain about
    assert fb.last_kernel() == kernel
    assert float(K.nzval.abs().max()) == 0.0 and close(f.cpu().numpy(), of)[0]
    with pytest.raises(fb.MissingPatternEntry):
        a = fb.start_assemble(K, f)
        fb.assemble_(a, full, cv)
        fb.finish_assemble(a)
    assert fb.last_kernel() == kernel
    # the same matrix entries as on the complete pattern wherever the entry exists
    ctx.synchronize()
    a = fb.start_assemble(K0, f)
    fb.assemble_(a, full, cv)
    fb.finish_assemble(a)
    assert close(K.nzval.cpu().numpy(), K0.nzval.cpu().numpy()[keep])[0]


@pytest.mark.parametrize("nel", [(3, 3, 3), (12, 10, 9)])   # per-cell kernel / tile kernel
def test_detj_not_positive_is_reported(ctx, nel):
    og = O.generate_grid("hexahedron", nel)
    cells = og.cells.copy()
    cells[7] = cells[7][[1, 0, 3, 2, 5, 4, 7, 6]]      # mirror one cell -> negative Jacobian
    g = fb.Grid.from_arrays(fb.Hexahedron, cells, og.nodes)
    ip = fb.Lagrange(fb.RefHexahedron, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    K = fb.allocate_matrix(dh)
    cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
    a = fb.start_assemble(K, ctx.zeros(dh.ndofs))
    a.variant = 5 if nel[0] > 10 else 0
    fb.assemble_(a, fb.HeatElement(), cv)
    with pytest.raises(fb.DetJNotPositive) as e:
        fb.finish_assemble(a)
    assert "cell 8" in str(e.value)


def test_goldens_through_the_gpu_path(ctx):
    # docs/src/topics/assembly.md:348-356
    g = fb.generate_grid(fb.Triangle, (100, 100))
    ip = fb.Lagrange(fb.RefTriangle, 2)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    assert dh.ndofs == 40401
    K = fb.allocate_matrix(dh)
    cv = fb.CellValues(fb.QuadratureRule(fb.RefTriangle, 2), ip)
    fb.assemble_(fb.start_assemble(K, None), fb.HeatElement(), cv)
    ref = 1138.8803468514259
    assert abs(float(K.nzval.norm()) - ref) / ref < 1e-13
    # docs/src/literate-tutorials/heat_equation.jl:59-114,181-234
    g = fb.generate_grid(fb.Quadrilateral, (20, 20))
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip)
    ch = fb.ConstraintHandler(dh)
    boundary = np.concatenate([fb.getfacetset(g, k) for k in ("left", "right", "top", "bottom")])
    fb.add_(ch, fb.Dirichlet("u", boundary, lambda x, t: 0))
    fb.close_(ch)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(), cv)
    fb.apply_(K, f, ch)
    u = spla.spsolve(K.tocsc(), f.cpu().numpy())
    ref = 3.307743912641305
    assert abs(np.linalg.norm(u) - ref) / ref < 1e-12


def test_full_size_properties_c2_scaled(ctx):
    """Size-independent properties at a large size the oracle cannot reach quickly: Laplace rows sum to
    zero, sum(f) = vol(Omega), K symmetric (config 2 at 96^3; bench.py checks the same at 200^3)."""
    import torch
    n = 96
    g = fb.generate_grid(fb.Hexahedron, (n, n, n)).perturb(0.2)
    ip = fb.Lagrange(fb.RefHexahedron, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    assert dh.ndofs == (n + 1) ** 3
    K = fb.allocate_matrix(dh)
    assert K.nnz == (3 * n + 1) ** 3
    f = ctx.zeros(dh.ndofs)
    cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(), cv)
    fb.finish_assemble(fb.start_assemble(K, f))
    assert abs(float(f.sum()) - 8.0) < 1e-11
    rows = torch.from_numpy(K.rowval - 1).to(f.device)
    rowsum = torch.zeros_like(f).index_add_(0, rows, K.nzval)
    assert float(rowsum.abs().max()) < 1e-12 * float(K.nzval.abs().max()) * 27
    # symmetry via x^T K y == y^T K x on random vectors
    cols = torch.repeat_interleave(torch.arange(K.n, device=f.device), torch.from_numpy(np.diff(K.colptr)).to(f.device))
    gen = torch.Generator(device=f.device).manual_seed(0)
    x = torch.rand(K.n, dtype=torch.float64, device=f.device, generator=gen)
    y = torch.rand(K.n, dtype=torch.float64, device=f.device, generator=gen)
    xKy = float((x[rows] * K.nzval * y[cols]).sum())
    yKx = float((y[rows] * K.nzval * x[cols]).sum())
    assert abs(xKy - yKx) <= 1e-11 * abs(xKy)


@pytest.mark.parametrize("mode", ["halo", "own"])
@pytest.mark.parametrize("ct,nel,order,vdim,qo,kind,p,nparts", [
    (fb.Hexahedron, (6, 5, 4), 1, 1, 2, "heat", {"k": 1.0, "source": 1.0}, 4),
    (fb.Hexahedron, (19, 11, 10), 1, 1, 2, "heat", {"k": 1.3, "source": 0.6}, 8),
    (fb.Hexahedron, (4, 4, 4), 1, 3, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}, 8),
    (fb.Hexahedron, (11, 9, 10), 1, 3, 2, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.3, 0.0, -1.0)}, 8),
    (fb.Hexahedron, (3, 3, 2), 2, 3, 3, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}, 2),
    (fb.Tetrahedron, (3, 2, 2), 2, 1, 2, "heat", {}, 3),
])
def test_partitioned_assembly_matches_serial_oracle(ctx, mode, ct, nel, order, vdim, qo, kind, p, nparts):
    """Multi-GPU path with all ranks emulated on one device: local problems, kernels, pack / unpack-add / mask are
    the product code; only the NCCL transport is replaced by a device-to-device hand-over (bench.py --gpus N and
    tests/test_partition_host.py cover the transport).  Gathered owned columns == serial oracle K, f."""
    import scipy.sparse as sp
    import torch
    hctx = fb.Context(-1)
    gg = fb.generate_grid(ct, nel, ctx=hctx).perturb(0.2)
    ip = fb.Lagrange(ct, order) ** vdim
    gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    elem, op = make_element(kind, p)
    oK = O.allocate_matrix(odh)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, kind, op)
    parts = [fb.Partition(gdh, nparts, r) for r in range(nparts)]
    st = []
    for pt in parts:
        lg, ldh = pt.local_problem(ctx)
        K = fb.allocate_matrix(ldh)
        f = ctx.zeros(ldh.ndofs)
        a = fb.start_assemble(K, f)
        pt.bind(a, cv)
        pt.assemble_(elem, mode=mode)
        if kind == "heat" and ct == fb.Hexahedron:      # block partition of a generated grid: structured view + cell map
            assert fb.last_kernel() == "k_march_hex"
        if kind == "elasticity" and ct == fb.Hexahedron and order == 1:
            assert fb.last_kernel() == "k_march_vec"
        st.append((lg, ldh, K, f, a))
    if mode == "own":
        for r, pt in enumerate(parts):
            for o in range(nparts):
                ns, fs, _, _ = pt.peer_counts(o)
                if ns + fs == 0:
                    continue
                buf = torch.empty(ns + fs, dtype=torch.float64, device=st[r][3].device)
                pt.pack(o, st[r][2], st[r][3], buf)
                parts[o].unpack_add(r, buf, st[o][2], st[o][3])
        for r, pt in enumerate(parts):
            pt.mask_unowned(st[r][2], st[r][3])
    rows, cols, vals, fd, fv = [], [], [], [], []
    for r, pt in enumerate(parts):
        i, j, v, d, x = pt.owned_triplets(st[r][2], st[r][3])
        rows.append(i); cols.append(j); vals.append(v); fd.append(d); fv.append(x)
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    fd, fv = np.concatenate(fd), np.concatenate(fv)
    assert len(np.unique(fd)) == odh.ndofs == len(fd)            # every dof owned exactly once
    # pattern bit-exact: gathered owned columns, sorted by (col, row), are exactly the oracle's CSC
    order_ = np.lexsort((rows, cols))
    assert len(rows) == oK.nnz
    assert np.array_equal(rows[order_], oK.rowval)
    assert np.array_equal(np.bincount(cols - 1, minlength=odh.ndofs), np.diff(oK.colptr))
    ok, nrm = close(vals[order_], oK.nzval)
    assert ok, nrm
    fg = np.zeros(odh.ndofs)
    fg[fd - 1] = fv
    ok, nrm = close(fg, of)
    assert ok, nrm


@pytest.mark.parametrize("nel,vdim,kind,p", [
    ((19, 11, 10), 1, "heat", {"k": 1.3, "source": 0.6}),
    ((21, 18, 13), 1, "heat", {"k": 1.0, "source": 1.0}),
    ((11, 9, 10), 3, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.3, 0.0, -1.0)}),
])
def test_split_marching_launch_of_the_exchange_path(ctx, nel, vdim, kind, p, monkeypatch):
    """exchange mode launches the CTAs with interface cells first and the others while the interface columns travel
    (fb2_assemble_distributed); FB2_MARCH_SPLIT=2 runs the same two launches without a communicator: every rank's local
    K, f must equal the single-launch result"""
    hctx = fb.Context(-1)
    ct = fb.Hexahedron
    gg = fb.generate_grid(ct, nel, ctx=hctx).perturb(0.2)
    ip = fb.Lagrange(ct, 1) ** vdim
    gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(ct, 2), ip)
    elem, _ = make_element(kind, p)
    for lz in ("", "2"):
        if lz:
            monkeypatch.setenv("FB2_MARCH_LZ", lz)
        else:
            monkeypatch.delenv("FB2_MARCH_LZ", raising=False)
        for r in range(8):
            pt = fb.Partition(gdh, 8, r)
            lg, ldh = pt.local_problem(ctx)
            K = fb.allocate_matrix(ldh)
            f = ctx.zeros(ldh.ndofs)
            a = fb.start_assemble(K, f)
            pt.bind(a, cv)
            monkeypatch.setenv("FB2_MARCH_SPLIT", "0")
            K.nzval.fill_(7.0)
            pt.assemble_(elem, mode="own")
            assert fb.last_kernel() == ("k_march_hex" if vdim == 1 else "k_march_vec")
            nz0, f0 = K.nzval.cpu().numpy().copy(), f.cpu().numpy().copy()
            monkeypatch.setenv("FB2_MARCH_SPLIT", "2")
            K.nzval.fill_(-5.0)
            f.fill_(3.0)
            pt.assemble_(elem, mode="own")
            assert close(K.nzval.cpu().numpy(), nz0, 1e-13)[0] and close(f.cpu().numpy(), f0, 1e-13)[0], (r, lz)


# ---- facet loop: FacetValues + Neumann / traction term ------------------------------------------------
@pytest.mark.parametrize("ct,nel,order,vdim,qo,kind,params,sets", [
    (fb.Hexahedron, (4, 3, 3), 1, 3, 2, "normal_traction", -0.1, ("top", "bottom", "front", "back")),
    (fb.Hexahedron, (3, 3, 2), 2, 3, 3, "traction", (0.3, -0.2, 0.7), ("right", "top")),
    (fb.Hexahedron, (3, 3, 3), 2, 1, 2, "flux", 2.5, ("left", "right", "top", "bottom", "front", "back")),
    (fb.Tetrahedron, (3, 3, 3), 1, 3, 1, "normal_traction", 0.4, ("top", "bottom", "front", "back", "left", "right")),
    (fb.Tetrahedron, (2, 3, 2), 2, 3, 3, "traction", (1.0, 2.0, -1.0), ("left", "back")),
    (fb.Quadrilateral, (5, 4), 2, 2, 3, "normal_traction", 1.5, ("left", "right", "top", "bottom")),
    (fb.Triangle, (4, 5), 2, 1, 2, "flux", -1.0, ("top", "left")),
    (fb.Triangle, (4, 3), 1, 2, 2, "traction", (0.5, 0.25), ("bottom", "right")),
])
def test_facet_loop_matches_oracle(ctx, ct, nel, order, vdim, qo, kind, params, sets):
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    ip, oip = cv.ip, ocv.ip
    fv = fb.FacetValues(fb.FacetQuadratureRule(ct, qo), ip)
    ofv = O.FacetValues(O.FacetQuadratureRule(SHAPE[ct], qo), oip)
    pairs = np.concatenate([fb.getfacetset(g, s) for s in sets])
    opairs = np.concatenate([np.asarray(og.facetsets[s]).reshape(-1, 2) for s in sets])
    assert np.array_equal(pairs, opairs)
    f = ctx.zeros(dh.ndofs)
    f += 1.0   # the facet loop adds onto f
    fb.assemble_facets_(f, dh, fv, pairs, kind, params)
    ctx.synchronize()
    of = np.ones(odh.ndofs)
    O.assemble_facets(odh, ofv, of, opairs, kind, params)
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, nrm


def test_hyperelasticity_tutorial_golden_through_the_gpu_path(ctx):
    # docs/src/literate-tutorials/hyperelasticity.jl:241-291,329-442: norm(u) == 4.761404305083876
    # K and g (Neo-Hooke volume terms + facet traction) and apply_zero! on the device, linear solves on the host
    import torch
    N, Lx = 10, 1.0
    g = fb.generate_grid(fb.Tetrahedron, (N, N, N), (0.0, 0.0, 0.0), (Lx, Lx, Lx))
    ip = fb.Lagrange(fb.RefTetrahedron, 1) ** 3
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.RefTetrahedron, 1), ip)
    fv = fb.FacetValues(fb.FacetQuadratureRule(fb.RefTetrahedron, 1), ip)
    th = np.pi / 3

    def rotation(x, t):
        return t * np.array([0.0, Lx / 2 - x[1] + (x[1] - Lx / 2) * np.cos(th) - (x[2] - Lx / 2) * np.sin(th),
                             Lx / 2 - x[2] + (x[1] - Lx / 2) * np.sin(th) + (x[2] - Lx / 2) * np.cos(th)])

    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), lambda x, t: [0.0, 0.0, 0.0], [1, 2, 3]))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), rotation, [1, 2, 3]))
    fb.close_(ch)
    fb.update_(ch, 0.5)
    gamma_n = fb.FacetSet(g, np.concatenate([fb.getfacetset(g, k) for k in ("top", "bottom", "front", "back")]))
    E, nu = 10.0, 0.3
    elem = fb.NeoHookeElement(lam=E * nu / ((1 + nu) * (1 - 2 * nu)), mu=E / (2 * (1 + nu)), b=(0.0, -0.5, 0.0))
    K = fb.allocate_matrix(dh)
    res = ctx.zeros(dh.ndofs)
    un = ctx.zeros(dh.ndofs)
    fb.apply_(un, ch)
    du = ctx.zeros(dh.ndofs)
    for it in range(31):
        u = un + du
        fb.assemble_(fb.start_assemble(K, res), elem, cv, u=u)
        fb.assemble_facets_(res, dh, fv, gamma_n, "normal_traction", -0.1)
        fb.apply_zero_(K, res, ch)
        if float(res.norm()) < 1e-8:
            break
        ddu = torch.from_numpy(spla.spsolve(K.tocsc(), res.cpu().numpy())).to(res.device)
        fb.apply_(ddu, ch, applyzero=True)
        du -= ddu
    else:
        raise AssertionError("Newton did not converge")
    ref = 4.761404305083876
    assert abs(float(u.norm()) - ref) / ref < 1e-7, float(u.norm())


# ---- the step after the path: SpMV, CSR values, CG on the device ------------------------------------
@pytest.mark.parametrize("ct,nel,order,vdim", [
    (fb.Hexahedron, (6, 5, 4), 1, 1), (fb.Hexahedron, (3, 3, 2), 2, 3), (fb.Tetrahedron, (3, 3, 3), 2, 1),
    (fb.Quadrilateral, (9, 7), 2, 2), (fb.Triangle, (8, 8), 1, 1), (fb.Hexahedron, (4, 4, 4), 1, 3),
])
def test_spmv_and_csr_values_match_scipy(ctx, ct, nel, order, vdim):
    import torch
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, 2)
    K = fb.allocate_matrix(dh)
    rng = np.random.default_rng(5)
    K.nzval.copy_(torch.from_numpy(rng.standard_normal(K.nnz)))     # non-symmetric values on the symmetric pattern
    x = torch.from_numpy(rng.standard_normal(K.n)).to(K.nzval.device)
    A = K.tocsc()
    y = fb.spmv(K, x).cpu().numpy()
    yt = fb.spmv(K, x, transpose=True).cpu().numpy()
    xr = x.cpu().numpy()
    assert close(y, A @ xr, 1e-13)[0]
    assert close(yt, A.T @ xr, 1e-13)[0]
    csr = A.tocsr()
    csr.sort_indices()
    assert np.array_equal(csr.indptr, K.colptr - 1) and np.array_equal(csr.indices, K.rowval - 1)   # rowptr = colptr, colval = rowval
    assert np.array_equal(fb.csr_values(K).cpu().numpy(), csr.data)


def test_spmv_on_a_non_symmetric_pattern(ctx):
    import torch
    g = fb.generate_grid(fb.Line, (4,))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefLine, 1)))
    n = dh.ndofs
    # upper bidiagonal pattern: entries (j-1, j) and (j, j)
    colptr = np.concatenate([[1], 1 + np.cumsum([1] + [2] * (n - 1))])
    rowval = np.concatenate([[1]] + [[j, j + 1] for j in range(1, n)])
    K = fb.allocate_matrix(dh, colptr, rowval)
    K.nzval.copy_(torch.arange(1.0, K.nnz + 1, dtype=torch.float64))
    x = torch.arange(1.0, n + 1, dtype=torch.float64, device=K.nzval.device)
    A = K.tocsc()
    assert close(fb.spmv(K, x).cpu().numpy(), A @ x.cpu().numpy(), 1e-14)[0]
    assert close(fb.spmv(K, x, transpose=True).cpu().numpy(), A.T @ x.cpu().numpy(), 1e-14)[0]
    with pytest.raises(fb.FB2Error):
        fb.csr_values(K)


def test_cg_solves_the_heat_tutorial(ctx):
    # heat_equation.jl:59-114,181-234 with the linear solve on the device: norm(u) == 3.307743912641305
    g = fb.generate_grid(fb.Quadrilateral, (20, 20))
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip)
    ch = fb.ConstraintHandler(dh)
    boundary = np.concatenate([fb.getfacetset(g, k) for k in ("left", "right", "top", "bottom")])
    fb.add_(ch, fb.Dirichlet("u", boundary, lambda x, t: 0))
    fb.close_(ch)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(), cv)
    fb.apply_(K, f, ch)
    for jacobi in (False, True):
        u = ctx.zeros(dh.ndofs)
        it, rn = fb.cg_(u, K, f, reltol=1e-13, jacobi=jacobi)
        assert 0 < it < dh.ndofs and rn <= 1e-13 * float(f.norm()) * 1.01
        ref = 3.307743912641305
        assert abs(float(u.norm()) - ref) / ref < 1e-11
    uref = spla.spsolve(K.tocsc(), f.cpu().numpy())
    assert close(u.cpu().numpy(), uref, 1e-10)[0]


def test_hyperelasticity_newton_cg_entirely_on_the_device(ctx):
    # hyperelasticity.jl:393-431: assemble, apply_zero!, cg!(ddu, K, g; maxiter = 1000), all on the device
    N, Lx = 10, 1.0
    g = fb.generate_grid(fb.Tetrahedron, (N, N, N), (0.0, 0.0, 0.0), (Lx, Lx, Lx))
    ip = fb.Lagrange(fb.RefTetrahedron, 1) ** 3
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.RefTetrahedron, 1), ip)
    fv = fb.FacetValues(fb.FacetQuadratureRule(fb.RefTetrahedron, 1), ip)
    th = np.pi / 3

    def rotation(x, t):
        return t * np.array([0.0, Lx / 2 - x[1] + (x[1] - Lx / 2) * np.cos(th) - (x[2] - Lx / 2) * np.sin(th),
                             Lx / 2 - x[2] + (x[1] - Lx / 2) * np.sin(th) + (x[2] - Lx / 2) * np.cos(th)])

    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), lambda x, t: [0.0, 0.0, 0.0], [1, 2, 3]))
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), rotation, [1, 2, 3]))
    fb.close_(ch)
    fb.update_(ch, 0.5)
    gamma_n = fb.FacetSet(g, np.concatenate([fb.getfacetset(g, k) for k in ("top", "bottom", "front", "back")]))
    E, nu = 10.0, 0.3
    elem = fb.NeoHookeElement(lam=E * nu / ((1 + nu) * (1 - 2 * nu)), mu=E / (2 * (1 + nu)), b=(0.0, -0.5, 0.0))
    K = fb.allocate_matrix(dh)
    res = ctx.zeros(dh.ndofs)
    un = ctx.zeros(dh.ndofs)
    fb.apply_(un, ch)
    du = ctx.zeros(dh.ndofs)
    for it in range(31):
        u = un + du
        fb.assemble_(fb.start_assemble(K, res), elem, cv, u=u)
        fb.assemble_facets_(res, dh, fv, gamma_n, "normal_traction", -0.1)
        fb.apply_zero_(K, res, ch)
        if float(res.norm()) < 1e-8:
            break
        ddu = ctx.zeros(dh.ndofs)
        fb.cg_(ddu, K, res, maxiter=1000)          # the tutorial's call: reltol = sqrt(eps)
        fb.apply_(ddu, ch, applyzero=True)
        du -= ddu
    else:
        raise AssertionError("Newton did not converge")
    ref = 4.761404305083876
    assert abs(float(u.norm()) - ref) / ref < 1e-7, float(u.norm())


@pytest.mark.parametrize("ct,nel,order,vdim,kind", [
    (fb.Hexahedron, (13, 11, 9), 1, 1, "heat"), (fb.Hexahedron, (5, 4, 6), 1, 3, "elasticity"),
    (fb.Quadrilateral, (40, 33), 2, 1, "heat"), (fb.Tetrahedron, (4, 4, 5), 2, 3, "neohooke"),
])
def test_streamed_host_path_matches_plain_host_path(ctx, ct, nel, order, vdim, kind):
    # fb2_assemble_host_streamed (slabs of cells, pipelined copies) == fb2_assemble_host, also after moving the nodes
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, 2 if order == 1 else 3)
    lam, mu = O.lame(10.0, 0.3)
    elem = {"heat": fb.HeatElement(1.3, 0.7), "elasticity": fb.ElasticityElement(lam=lam, mu=mu, b=(0.1, 0.2, -1.0)),
            "neohooke": fb.NeoHookeElement(lam=lam, mu=mu, b=(0.0, -0.5, 0.0))}[kind]
    K = fb.allocate_matrix(dh)
    u = 0.01 * np.sin(np.arange(dh.ndofs, dtype=np.float64)) if kind == "neohooke" else None
    nz1, f1 = np.empty(K.nnz), np.empty(dh.ndofs)
    nz2, f2 = np.full(K.nnz, np.nan), np.full(dh.ndofs, np.nan)
    fb.assemble_host(fb.start_assemble(K, None), elem, cv, nz1, f1, u=u)
    fb.assemble_host_streamed(fb.start_assemble(K, None), elem, cv, nz2, f2, u=u)
    assert np.array_equal(np.isnan(nz2), np.zeros(K.nnz, bool))
    assert close(nz2, nz1, 1e-13)[0] and close(f2, f1, 1e-13)[0]
    # new coordinates travel with the call
    xyz = np.ascontiguousarray(g.nodes * 1.25)
    fb.assemble_host_streamed(fb.start_assemble(K, None), elem, cv, nz2, f2, u=u, xyz=xyz)
    g.upload_coordinates_async(xyz)
    fb.assemble_host(fb.start_assemble(K, None), elem, cv, nz1, f1, u=u)
    assert close(nz2, nz1, 1e-13)[0] and close(f2, f1, 1e-13)[0]


# ---- property test: random small problems on the device against the oracle ------------------------------------------
from hypothesis import HealthCheck, given, settings, strategies as hst  # noqa: E402


@hst.composite
def _random_case(draw):
    ct = draw(hst.sampled_from([fb.Triangle, fb.Quadrilateral, fb.Tetrahedron, fb.Hexahedron]))
    dim = 2 if ct in (fb.Triangle, fb.Quadrilateral) else 3
    nel = tuple(draw(hst.integers(1, 5 if dim == 3 else 9)) for _ in range(dim))
    order = draw(hst.integers(1, 2))
    kind = draw(hst.sampled_from(["heat", "mass", "elasticity"] + (["neohooke"] if dim == 3 else [])))
    vdim = 1 if kind in ("heat", "mass") else dim
    qo = draw(hst.integers(max(1, order), 3 if ct == fb.Triangle else 4)) if kind != "neohooke" else 2
    scatter = draw(hst.sampled_from(["atomic", "colored"]))
    return ct, nel, order, vdim, qo, kind, scatter


_PROP_N = int(os.environ.get("FB2_PROP_EXAMPLES", "0"))      # > 0: a longer, randomly seeded exploration run


@settings(max_examples=_PROP_N or 40, deadline=None, derandomize=_PROP_N == 0,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(_random_case())
def test_random_small_problems_match_oracle(ctx, case):
    import torch
    ct, nel, order, vdim, qo, kind, scatter = case
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    p = {"heat": {"k": 1.7, "source": 0.3}, "mass": {"rho": 2.5}}.get(kind, {"E": 10.0, "nu": 0.3, "b": (0.1, -0.5, 0.2)[:og.sdim]})
    elem, op = make_element(kind, p)
    K = fb.allocate_matrix(dh)
    oK = O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f = ctx.zeros(dh.ndofs)
    of = np.zeros(odh.ndofs)
    u = ou = None
    if kind == "neohooke":
        ou = 0.2 * displacement(og, odh, vdim)
        u = torch.from_numpy(ou).to(f.device)
    fb.assemble_(fb.start_assemble(K, f, scatter=scatter), elem, cv, u=u)
    ctx.synchronize()
    O.assemble_global(odh, ocv, oK, of, kind, op, u=ou)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, (case, nrm)
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, (case, nrm)


@pytest.mark.parametrize("ct,nel,order,qo", [
    (fb.Hexahedron, (4, 3, 5), 1, 2), (fb.Hexahedron, (3, 3, 2), 2, 3), (fb.Tetrahedron, (3, 2, 2), 2, 4),
    (fb.Quadrilateral, (6, 5), 2, 3), (fb.Triangle, (5, 4), 1, 2),
])
def test_standalone_reinit_matches_oracle(ctx, ct, nel, order, qo):
    # reinit!(cv, cell): dNdx = shape_gradient(cv, q, i), detJdV = getdetJdV(cv, q) (src/FEValues/CellValues.jl:122-140)
    g, og, dh, odh, cv, ocv = build(ct, nel, order, 1, qo)
    dNdx, dO = fb.reinit_(cv, g)
    odN, odO = O.reinit(ocv, og.nodes[og.cells - 1])
    assert close(dNdx.cpu().numpy(), odN, 1e-13)[0]
    assert close(dO.cpu().numpy(), odO, 1e-13)[0]
    ids = np.array([og.ncells, 1, 2], dtype=np.int64)
    dNdx, dO = fb.reinit_(cv, g, ids)
    assert close(dNdx.cpu().numpy(), odN[ids - 1], 1e-13)[0] and close(dO.cpu().numpy(), odO[ids - 1], 1e-13)[0]


@pytest.mark.parametrize("ct,nel,order,vdim,qo", [
    (fb.Hexahedron, (4, 3, 5), 1, 3, 2), (fb.Hexahedron, (3, 3, 2), 2, 1, 3), (fb.Tetrahedron, (3, 2, 2), 2, 3, 4),
    (fb.Quadrilateral, (6, 5), 2, 2, 3), (fb.Triangle, (5, 4), 1, 1, 2),
])
def test_function_values_and_gradients_match_oracle(ctx, ct, nel, order, vdim, qo):
    # function_value / function_gradient (src/FEValues/common_values.jl:177-227) at every quadrature point
    import torch
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    rng = np.random.default_rng(11)
    u = rng.standard_normal(odh.ndofs)
    vals, grads = fb.function_values_(cv, dh, torch.from_numpy(u).to(f"cuda:{ctx.device}"))
    dNdx, _ = O.reinit(ocv, og.nodes[og.cells - 1])                 # (nc, nq, nb, dim)
    ue = u[odh.cell_dofs - 1].reshape(og.ncells, -1, vdim)           # (nc, nb, vdim)
    oval = np.einsum("qa,cav->cqv", ocv.N, ue)
    ograd = np.einsum("cav,cqad->cqvd", ue, dNdx)
    assert close(vals.cpu().numpy(), oval, 1e-13)[0]
    assert close(grads.cpu().numpy(), ograd, 1e-13)[0]


@hst.composite
def _random_facet_case(draw):
    ct = draw(hst.sampled_from([fb.Triangle, fb.Quadrilateral, fb.Tetrahedron, fb.Hexahedron]))
    dim = 2 if ct in (fb.Triangle, fb.Quadrilateral) else 3
    nel = tuple(draw(hst.integers(1, 4 if dim == 3 else 7)) for _ in range(dim))
    order = draw(hst.integers(1, 2))
    kind = draw(hst.sampled_from(["flux", "traction", "normal_traction"]))
    vdim = 1 if kind == "flux" else dim
    qo = draw(hst.integers(1, 3))
    names = ["left", "right", "top", "bottom"] + (["front", "back"] if dim == 3 else [])
    sets = draw(hst.lists(hst.sampled_from(names), min_size=1, max_size=len(names), unique=True))
    return ct, nel, order, vdim, qo, kind, tuple(sets)


@settings(max_examples=_PROP_N or 30, deadline=None, derandomize=_PROP_N == 0,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(_random_facet_case())
def test_random_facet_loops_match_oracle(ctx, case):
    ct, nel, order, vdim, qo, kind, sets = case
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, max(qo, 1))
    fv = fb.FacetValues(fb.FacetQuadratureRule(ct, qo), cv.ip)
    ofv = O.FacetValues(O.FacetQuadratureRule(SHAPE[ct], qo), ocv.ip)
    pairs = np.concatenate([fb.getfacetset(g, s) for s in sets])
    params = {"flux": 1.75, "traction": (0.3, -0.2, 0.7)[:og.sdim], "normal_traction": -0.4}[kind]
    f = ctx.zeros(dh.ndofs)
    fb.assemble_facets_(f, dh, fv, pairs, kind, params)
    ctx.synchronize()
    of = np.zeros(odh.ndofs)
    O.assemble_facets(odh, ofv, of, pairs, kind, params)
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, (case, nrm)


@pytest.mark.parametrize("ct,nel,order,vdim,qo,kind,p,nparts", [
    (fb.Hexahedron, (5, 4, 3), 1, 1, 2, "heat", {"k": 1.0, "source": 1.0}, 3),
    (fb.Tetrahedron, (3, 2, 2), 1, 3, 2, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.0, 0.0, -1.0)}, 4),
])
@pytest.mark.parametrize("how", ["random_owners", "metis"])
def test_partition_from_random_owner_array_matches_serial_oracle(ctx, ct, nel, order, vdim, qo, kind, p, nparts, how):
    """A scattered (worst-case) cell -> rank array, as an external partitioner would hand in, and METIS_PartMeshDual
    inside the library: gathered owned columns of the emulated ranks == serial oracle."""
    import torch
    hctx = fb.Context(-1)
    gg = fb.generate_grid(ct, nel, ctx=hctx).perturb(0.2)
    ip = fb.Lagrange(ct, order) ** vdim
    gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    elem, op = make_element(kind, p)
    oK = O.allocate_matrix(odh)
    of = np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, kind, op)
    owner = np.random.default_rng(3).integers(0, nparts, size=gg.ncells).astype(np.int32)
    owner[:nparts] = np.arange(nparts)
    if how == "metis":
        parts = [fb.Partition(gdh, nparts, r, metis=True) for r in range(nparts)]
    else:
        parts = [fb.Partition(gdh, nparts, r, cell_owner=owner) for r in range(nparts)]
    st = []
    for pt in parts:
        lg, ldh = pt.local_problem(ctx)
        K = fb.allocate_matrix(ldh)
        f = ctx.zeros(ldh.ndofs)
        a = fb.start_assemble(K, f)
        pt.bind(a, cv)
        pt.assemble_(elem, mode="own")
        st.append((lg, ldh, K, f, a))
    for r, pt in enumerate(parts):
        for o in range(nparts):
            ns, fs, _, _ = pt.peer_counts(o)
            if ns + fs == 0:
                continue
            buf = torch.empty(ns + fs, dtype=torch.float64, device=st[r][3].device)
            pt.pack(o, st[r][2], st[r][3], buf)
            parts[o].unpack_add(r, buf, st[o][2], st[o][3])
    for r, pt in enumerate(parts):
        pt.mask_unowned(st[r][2], st[r][3])
    rows, cols, vals, fd, fv = [], [], [], [], []
    for r, pt in enumerate(parts):
        i, j, v, d, x = pt.owned_triplets(st[r][2], st[r][3])
        rows.append(i); cols.append(j); vals.append(v); fd.append(d); fv.append(x)
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    fd, fv = np.concatenate(fd), np.concatenate(fv)
    order_ = np.lexsort((rows, cols))
    assert len(rows) == oK.nnz and np.array_equal(rows[order_], oK.rowval)
    ok, nrm = close(vals[order_], oK.nzval)
    assert ok, nrm
    fg = np.zeros(odh.ndofs)
    fg[fd - 1] = fv
    ok, nrm = close(fg, of)
    assert ok, nrm


def test_empty_inputs_are_no_ops(ctx):
    # empty facet set, empty Dirichlet set, reinit of zero cells: nothing changes, nothing fails
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, (2, 2, 2), 1, 1, 2)
    fv = fb.FacetValues(fb.FacetQuadratureRule(fb.Hexahedron, 2), cv.ip)
    f = ctx.zeros(dh.ndofs)
    f += 3.0
    fb.assemble_facets_(f, dh, fv, np.zeros((0, 2), dtype=np.int64), "flux", 1.0)
    ctx.synchronize()
    assert float((f - 3.0).abs().max()) == 0.0
    dNdx, dO = fb.reinit_(cv, g, np.zeros(0, dtype=np.int64))
    assert dNdx.shape[0] == 0 and dO.shape[0] == 0
    K = fb.allocate_matrix(dh)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(), cv)
    ref = K.nzval.clone()
    ch = fb.ConstraintHandler.from_arrays(dh, np.zeros(0, dtype=np.int64), np.zeros(0))
    fb.apply_(K, f, ch)
    ctx.synchronize()
    assert float((K.nzval - ref).abs().max()) == 0.0


@pytest.mark.parametrize("ct,nel,order,vdim,qo,kind,p", [
    (fb.Hexahedron, (5, 4, 3), 1, 3, 2, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    (fb.Hexahedron, (14, 12, 11), 1, 1, 2, "heat", {"k": 2.0, "source": 3.0}),      # tile kernel
    (fb.Tetrahedron, (3, 3, 2), 2, 3, 4, "elasticity", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
])
@pytest.mark.parametrize("order_kind", ["componentwise", "permutation"])
def test_renumbered_problem_matches_oracle(ctx, ct, nel, order, vdim, qo, kind, p, order_kind):
    # renumber!(dh, ch, order) (src/Dofs/DofRenumbering.jl:79-125) followed by pattern, assembly, update! and apply! on the
    # device; the oracle renumbers its DofHandler first and builds its ConstraintHandler from the new numbering
    g, og, dh, odh, cv, ocv = build(ct, nel, order, vdim, qo)
    elem, op = make_element(kind, p)

    def val(x, t):
        return [0.01 * x[1] + 0.02 * k + t for k in range(vdim)] if vdim > 1 else 0.3 * x[0] - x[1] + t
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), val))
    fb.close_(ch)
    if order_kind == "componentwise":
        perm = fb.renumber_(dh, ch, fb.DofOrder.ComponentWise())
        assert np.array_equal(perm, O.renumber_permutation(odh, "componentwise"))
    else:
        perm = fb.renumber_(dh, ch, np.random.default_rng(5).permutation(dh.ndofs) + 1)
    O.renumber(odh, perm)
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    och = O.ConstraintHandler(odh)
    och.add(O.Dirichlet("u", og.facetsets["left"], val))
    och.close()
    fb.update_(ch, 0.25)
    och.update(0.25)
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)
    K, oK = fb.allocate_matrix(dh), O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f, of = ctx.zeros(dh.ndofs), np.zeros(odh.ndofs)
    O.assemble_global(odh, ocv, oK, of, kind, op)
    for scatter in ("atomic", "colored"):
        fb.assemble_(fb.start_assemble(K, f, scatter=scatter), elem, cv)
        ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
        assert ok, (scatter, nrm)
        ok, nrm = close(f.cpu().numpy(), of)
        assert ok, (scatter, nrm)
    fb.apply_(K, f, ch)
    och.apply(oK, of)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, nrm


def test_config_c1_quad_100x100_entrywise_with_apply(ctx):
    """BASELINE.json configs[0] at its full size: heat equation, Q1 on generate_grid(Quadrilateral, (100, 100)) with the
    tutorial's homogeneous Dirichlet condition on all four facet sets (heat_equation.jl:59-114): numbering and pattern
    bit-exact, K / f entry-wise before and after apply!, and the solution of the constrained system."""
    nel = (100, 100)
    g, og, dh, odh, cv, ocv = build(fb.Quadrilateral, nel, 1, 1, 2, perturb=False)
    assert g.ncells == 10000 and dh.ndofs == 10201
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    K, oK = fb.allocate_matrix(dh), O.allocate_matrix(odh)
    assert K.nnz == 90601
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f, of = ctx.zeros(dh.ndofs), np.zeros(odh.ndofs)
    fb.assemble_(fb.start_assemble(K, f), fb.HeatElement(), cv)
    O.assemble_global(odh, ocv, oK, of, "heat")
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    ch, och = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    union = np.concatenate([fb.getfacetset(g, s) for s in ("left", "right", "top", "bottom")])
    ounion = np.concatenate([og.facetsets[s] for s in ("left", "right", "top", "bottom")])
    fb.add_(ch, fb.Dirichlet("u", union, lambda x, t: 0.0))
    och.add(O.Dirichlet("u", ounion, lambda x, t: 0.0))
    fb.close_(ch)
    och.close()
    fb.update_(ch, 0.0)
    och.update(0.0)
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs) and len(ch.prescribed_dofs) == 400
    fb.apply_(K, f, ch)
    och.apply(oK, of)
    assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    u = spla.spsolve(K.tocsc(), f.cpu().numpy())
    ou = spla.spsolve(oK.toscipy().tocsc(), of)
    assert np.linalg.norm(u - ou) <= 1e-10 * np.linalg.norm(ou)


@pytest.mark.parametrize("kind,nel,vdim", [("heat", (64, 64, 64), 1), ("elasticity", (32, 32, 32), 3), ("heat", (96, 40, 72), 1),
                                           ("elasticity", (45, 38, 41), 3)])
def test_large_entrywise_against_the_c_port(ctx, kind, nel, vdim):
    """Entry-wise nzval / f at sizes the numpy oracle does not reach: the C restatement of the reference loop
    (oracle/cpu_assemble.c, itself checked against the numpy oracle in tests/test_oracle_goldens.py) is the checker.
    64^3 Q1 heat (262 144 cells: many tiles, chunks and waves of the marching kernel) and 32^3 Q1^3 elasticity."""
    from oracle import cport
    g, og, dh, odh, cv, ocv = build(fb.Hexahedron, nel, 1, vdim, 2, True)
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    K, oK = fb.allocate_matrix(dh), O.allocate_matrix(odh)
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
    f, of = ctx.zeros(dh.ndofs), np.zeros(odh.ndofs)
    if kind == "heat":
        elem, op = fb.HeatElement(k=1.3, source=0.7), dict(k=1.3, source=0.7)
    else:
        lam, mu = O.lame(200e9, 0.3)
        elem, op = fb.ElasticityElement(lam=lam, mu=mu, b=(0.1, 0.2, -1.0)), {"lambda": lam, "mu": mu, "b": (0.1, 0.2, -1.0)}
    cport.assemble(odh, ocv, oK, of, kind, op, nthreads=1)      # one thread: the reference's serial summation order
    fb.assemble_(fb.start_assemble(K, f), elem, cv)
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm
    ok, nrm = close(f.cpu().numpy(), of)
    assert ok, nrm


@pytest.mark.parametrize("ct,nel,order,qo", [(fb.Hexahedron, (5, 4, 3), 1, 2), (fb.Hexahedron, (13, 9, 10), 1, 2), (fb.Hexahedron, (3, 2, 2), 2, 3),
                                              (fb.Tetrahedron, (3, 3, 2), 2, 4), (fb.Quadrilateral, (7, 5), 1, 2), (fb.Triangle, (5, 4), 2, 3)])
def test_general_stiffness_tensor_elasticity(ctx, ct, nel, order, qo):
    """FB2_ELEM_ELASTICITY_GENERAL: any SymmetricTensor{4} C (linear_elasticity.jl:266-281, benchmark/helper.jl:249-262).  An
    orthotropic C against the oracle's einsum restatement, and the isotropic C must reproduce the tensor-core SYRK element."""
    dim = len(nel)
    g, og, dh, odh, cv, ocv = build(ct, nel, order, dim, qo)
    rng = np.random.default_rng(3)
    # orthotropic: random symmetric positive definite 6x6 (3x3 in 2-D) Voigt matrix with the orthotropic zero pattern
    voigt = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)] if dim == 3 else [(0, 0), (1, 1), (0, 1)]
    nv = len(voigt)
    D = np.zeros((nv, nv))
    A = rng.random((dim, dim)) + dim * np.eye(dim)
    D[:dim, :dim] = A @ A.T
    D[dim:, dim:] = np.diag(rng.random(nv - dim) + 0.5)
    C4 = np.zeros((dim,) * 4)
    for I, (i, j) in enumerate(voigt):
        for J, (k, l) in enumerate(voigt):
            for (a, b_) in {(i, j), (j, i)}:
                for (c, d) in {(k, l), (l, k)}:
                    C4[a, b_, c, d] = D[I, J]
    bf = (0.3, -1.0, 0.5)[:dim]
    K, oK = fb.allocate_matrix(dh), O.allocate_matrix(odh)
    f, of = ctx.zeros(dh.ndofs), np.zeros(odh.ndofs)
    fb.assemble_(fb.start_assemble(K, f), fb.GeneralElasticityElement(C4, bf), cv)
    if ct == fb.Hexahedron and order == 1:       # structured trilinear hexahedra: the marching-tile kernel takes C as an argument
        assert fb.last_kernel() == "k_march_vec"
    O.assemble_global(odh, ocv, oK, of, "elasticity_general", {"C": C4, "b": bf})
    ok, nrm = close(K.nzval.cpu().numpy(), oK.nzval)
    assert ok, nrm
    assert close(f.cpu().numpy(), of)[0]
    if ct == fb.Hexahedron and order == 1:       # and the block kernel gives the same matrix
        a1 = fb.start_assemble(K, f)
        a1.variant = 1
        fb.assemble_(a1, fb.GeneralElasticityElement(C4, bf), cv)
        assert fb.last_kernel() == "k_cell_blocks"
        assert close(K.nzval.cpu().numpy(), oK.nzval)[0] and close(f.cpu().numpy(), of)[0]
    # isotropic C == ElasticityElement
    lam, mu = O.lame(10.0, 0.3)
    fb.assemble_(fb.start_assemble(K, f), fb.GeneralElasticityElement(O.isotropic_stiffness(lam, mu, dim), bf), cv)
    nz_general = K.nzval.cpu().numpy().copy()
    fb.assemble_(fb.start_assemble(K, f), fb.ElasticityElement(lam=lam, mu=mu, b=bf), cv)
    assert close(nz_general, K.nzval.cpu().numpy(), 1e-12)[0]
    with pytest.raises(fb.FB2Error, match="minor symmetries"):
        bad = C4.copy()
        bad[0, 1, 0, 0] += 1.0
        fb.assemble_(fb.start_assemble(K, f), fb.GeneralElasticityElement(bad, bf), cv)

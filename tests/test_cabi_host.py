"""CPU-side tests of the C ABI (no GPU): the library loads, exports every declared symbol, fails loudly
without a device, and its host logic (grid generation, dof numbering, Dirichlet set-up, quadrature and
shape tables) is bit-identical to the oracle / the reference's literal goldens."""
import numpy as np
import pytest

import ferrite_b200 as fb
import oracle as O

SHAPE = {fb.Line: "line", fb.Triangle: "triangle", fb.Quadrilateral: "quadrilateral",
         fb.Tetrahedron: "tetrahedron", fb.Hexahedron: "hexahedron"}


@pytest.fixture(scope="module")
def hctx():
    return fb.Context(-1)


def test_library_exports_every_declared_symbol():
    syms = fb.declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(fb.lib, s)]
    assert missing == []


def test_no_cpu_fallback(hctx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.FB2Error) as e:
        fb.Context(0, use_torch_stream=False)
    assert "no CPU fallback" in str(e.value)
    g = fb.generate_grid(fb.Quadrilateral, (2, 2), ctx=hctx)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefQuadrilateral, 1)))
    with pytest.raises(fb.FB2Error) as e:
        fb.allocate_matrix(dh)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("ct,nel,left,right", [
    (fb.Line, (5,), None, None),
    (fb.Quadrilateral, (4, 3), None, None),
    (fb.Quadrilateral, (7, 5), (0.0, -2.0), (3.0, 0.5)),
    (fb.Triangle, (3, 4), None, None),
    (fb.Hexahedron, (3, 2, 4), None, None),
    (fb.Hexahedron, (5, 4, 3), (0.1, 0.2, 0.3), (1.7, 0.9, 2.2)),
    (fb.Tetrahedron, (2, 3, 2), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
])
def test_generate_grid_matches_oracle(hctx, ct, nel, left, right):
    g = fb.generate_grid(ct, nel, left, right, ctx=hctx)
    og = O.generate_grid(SHAPE[ct], nel, left, right)
    assert np.array_equal(g.cells, og.cells)
    assert np.array_equal(g.nodes, og.nodes)          # same formula, same operation order: bit-identical
    for name, pairs in og.facetsets.items():
        assert np.array_equal(fb.getfacetset(g, name), pairs), name


@pytest.mark.parametrize("ct,nel", [(fb.Hexahedron, (23, 17, 9)), (fb.Hexahedron, (2, 1, 3)), (fb.Tetrahedron, (5, 4, 3)), (fb.Quadrilateral, (31, 7))])
def test_host_setup_does_not_depend_on_the_thread_count(hctx, ct, nel, monkeypatch):
    """the independent host loops of the set-up (nodes, connectivity, perturbation, partition arrays) run on several threads
    (FB2_HOST_THREADS): grid, facet sets and partition plan must be identical for 1 and 5 threads"""
    out = []
    for threads in ("1", "5"):
        monkeypatch.setenv("FB2_HOST_THREADS", threads)
        g = fb.generate_grid(ct, nel, ctx=hctx).perturb(0.15)
        dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, 1)))
        part = fb.Partition(dh, 4, 1)
        out.append((g.cells.copy(), g.nodes.copy(), {k: np.asarray(fb.getfacetset(g, k)).copy() for k in ("left", "right")},
                    part.l2g_dof.copy(), part.dof_owner.copy(), part.cells_global.copy(),
                    (part.ncells_local, part.ncells_own, part.nnodes_local, part.ndofs_local, part.ndofs_owned)))
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert all(np.array_equal(a[2][k], b[2][k]) for k in a[2])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5]) and a[6] == b[6]


def test_perturb_matches_oracle(hctx):
    nel, left, right = (4, 3, 5), (0.0, 0.0, 0.0), (1.0, 2.0, 3.0)
    g = fb.generate_grid(fb.Hexahedron, nel, left, right, ctx=hctx).perturb(0.2)
    og = O.perturb_grid(O.generate_grid("hexahedron", nel, left, right), nel, left, right, 0.2)
    assert np.array_equal(g.nodes, og.nodes)
    assert not np.array_equal(og.nodes, O.generate_grid("hexahedron", nel, left, right).nodes)


FIELDSETS = [
    (fb.Quadrilateral, (5, 4), [("u", 1, 1)]),
    (fb.Quadrilateral, (5, 4), [("u", 2, 1)]),
    (fb.Quadrilateral, (3, 3), [("v", 2, 2), ("s", 1, 1)]),
    (fb.Triangle, (4, 3), [("u", 2, 2), ("p", 1, 1)]),
    (fb.Hexahedron, (3, 3, 2), [("u", 1, 1)]),
    (fb.Hexahedron, (3, 2, 2), [("u", 1, 3)]),
    (fb.Hexahedron, (3, 2, 3), [("u", 2, 3)]),
    (fb.Hexahedron, (2, 2, 2), [("u", 2, 3), ("p", 1, 1)]),
    (fb.Tetrahedron, (2, 2, 3), [("u", 1, 3)]),
    (fb.Tetrahedron, (3, 2, 2), [("u", 2, 3)]),
    (fb.Tetrahedron, (2, 2, 2), [("u", 2, 3), ("p", 1, 1)]),
]


def _both_dh(hctx, ct, nel, fields):
    g = fb.generate_grid(ct, nel, ctx=hctx)
    dh = fb.DofHandler(g)
    og = O.generate_grid(SHAPE[ct], nel)
    odh = O.DofHandler(og)
    for name, order, vdim in fields:
        fb.add_(dh, name, fb.Lagrange(ct, order) ** vdim)
        odh.add(name, O.Lagrange(SHAPE[ct], order) ** vdim if vdim > 1 else O.Lagrange(SHAPE[ct], order))
    fb.close_(dh)
    odh.close()
    return g, dh, og, odh


@pytest.mark.parametrize("ct,nel,fields", FIELDSETS)
def test_dof_numbering_bit_exact(hctx, ct, nel, fields):
    g, dh, og, odh = _both_dh(hctx, ct, nel, fields)
    assert dh.ndofs == odh.ndofs
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs)
    for name, _, _ in fields:
        r = fb.dof_range(dh, name)
        assert (r[0], r[-1]) == odh.dof_range(name)


def test_dof_goldens_through_cabi(hctx):
    # test/test_dofs.jl:95-126 (shell quads) and :235-257 (two fields on 2x1 quads)
    nodes = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [2, 2, 0]], dtype=float)
    g = fb.Grid.from_arrays(fb.Quadrilateral, [(1, 2, 3, 4), (2, 5, 6, 3)], nodes, ctx=hctx)
    q1, q2 = fb.Lagrange(fb.RefQuadrilateral, 1), fb.Lagrange(fb.RefQuadrilateral, 2)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", q1 ** 3), "th", q1 ** 3))
    assert list(fb.celldofs(dh, 1)) == list(range(1, 25))
    assert list(fb.celldofs(dh, 2)) == [4, 5, 6, 25, 26, 27, 28, 29, 30, 7, 8, 9, 16, 17, 18, 31, 32, 33, 34, 35, 36, 19, 20, 21]
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", q2), "th", q2))
    assert list(fb.celldofs(dh, 1)) == list(range(1, 19))
    assert list(fb.celldofs(dh, 2)) == [2, 19, 20, 3, 21, 22, 23, 6, 24, 11, 25, 26, 12, 27, 28, 29, 15, 30]
    # edge bc on the shell: test/test_constraints.jl:175-200
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("th", [(1, 1), (2, 1)], lambda x, t: 0.0, [1], kind="edge"))
    fb.close_(ch)
    assert list(ch.prescribed_dofs) == [10, 11, 14, 25, 27]

    g = fb.generate_grid(fb.Quadrilateral, (2, 1), ctx=hctx)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "v", q1 ** 2), "s", q1))
    assert list(fb.celldofs(dh, 1)) == list(range(1, 13))
    assert list(fb.celldofs(dh, 2)) == [3, 4, 13, 14, 15, 16, 5, 6, 10, 17, 18, 11]
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("v", fb.getfacetset(g, "left"), lambda x, t: 0, [2]))
    fb.add_(ch, fb.Dirichlet("s", fb.getfacetset(g, "left"), lambda x, t: 0))
    fb.close_(ch)
    assert list(ch.prescribed_dofs) == [2, 8, 9, 12]


def test_constraint_goldens_through_cabi(hctx):
    # test/test_constraints.jl:84-101 (node sets) and :155-172 (edge set on a hex)
    g = fb.generate_grid(fb.Triangle, (1, 1), ctx=hctx)
    nodeset = [i + 1 for i, x in enumerate(g.nodes) if x[1] == -1 or x[0] == -1]
    p1 = fb.Lagrange(fb.RefTriangle, 1)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", p1 ** 2), "p", p1))
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", nodeset, lambda x, t: x, [1, 2], kind="node"))
    fb.add_(ch, fb.Dirichlet("p", nodeset, lambda x, t: 0, 1, kind="node"))
    fb.close_(ch)
    assert list(ch.prescribed_dofs) == list(range(1, 10))
    assert list(ch.inhomogeneities) == [-1, -1, 1, -1, -1, 1, 0, 0, 0]

    g = fb.generate_grid(fb.Hexahedron, (1, 1, 1), ctx=hctx)
    h1 = fb.Lagrange(fb.RefHexahedron, 1)
    dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "u", h1 ** 3), "p", h1))
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet("u", [(1, 4)], lambda x, t: x, [1, 2, 3], kind="edge"))
    fb.close_(ch)
    assert list(ch.prescribed_dofs) == [1, 2, 3, 10, 11, 12]
    assert list(ch.inhomogeneities) == [-1.0, -1.0, -1.0, -1.0, 1.0, -1.0]


@pytest.mark.parametrize("ct,nel,fields", [f for f in FIELDSETS if f[0] != fb.Line])
def test_dirichlet_matches_oracle(hctx, ct, nel, fields):
    g, dh, og, odh = _both_dh(hctx, ct, nel, fields)
    ch, och = fb.ConstraintHandler(dh), O.ConstraintHandler(odh)
    names = sorted(og.facetsets)
    name, order, vdim = fields[0]

    def f1(x, t):
        return [np.sin(x[0] + 0.3 * k) + t + x[-1] for k in range(vdim)]

    def f2(x, t):
        return 0.25 * x[0] - x[1]
    fb.add_(ch, fb.Dirichlet(name, fb.getfacetset(g, names[0]), f1))
    och.add(O.Dirichlet(name, og.facetsets[names[0]], f1))
    # overlapping second condition on one component: later conditions override earlier ones
    both = np.concatenate([og.facetsets[names[0]][:2], og.facetsets[names[1]]])
    fb.add_(ch, fb.Dirichlet(name, both, f2, [1]))
    och.add(O.Dirichlet(name, both, f2, [1]))
    fb.close_(ch)
    och.close()
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)
    fb.update_(ch, 1.5)
    och.update(1.5)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)


@pytest.mark.parametrize("ct,qo,io,vdim", [
    (fb.Quadrilateral, 2, 1, 1), (fb.Quadrilateral, 3, 2, 2), (fb.Triangle, 2, 2, 1), (fb.Triangle, 1, 1, 2),
    (fb.Hexahedron, 2, 1, 1), (fb.Hexahedron, 3, 2, 3), (fb.Tetrahedron, 2, 1, 3), (fb.Tetrahedron, 4, 2, 3),
    (fb.Tetrahedron, 3, 2, 1), (fb.Hexahedron, 4, 2, 1),
])
def test_cellvalues_tables_match_oracle(hctx, ct, qo, io, vdim):
    cv = fb.CellValues(fb.QuadratureRule(ct, qo), fb.Lagrange(ct, io) ** vdim, ctx=hctx)
    t = cv.tables()
    oip = O.Lagrange(SHAPE[ct], io)
    ocv = O.CellValues(O.QuadratureRule(SHAPE[ct], qo), oip ** vdim if vdim > 1 else oip)
    assert cv.nq == ocv.nq and cv.nbase_scalar == oip.nbase
    # Gauss points come from two independent generators (Newton vs numpy's eigen-solve): ~1 ulp apart
    assert np.allclose(t["w"], ocv.w, rtol=1e-14, atol=0)
    assert np.allclose(t["points"], ocv.qr.points, rtol=1e-14, atol=1e-16)
    assert np.allclose(t["N"], ocv.N, rtol=1e-13, atol=1e-15)
    assert np.allclose(t["dNdxi"], ocv.dNdxi, rtol=1e-13, atol=1e-15)
    assert np.allclose(t["M"], ocv.M, rtol=1e-13, atol=1e-15)
    assert np.allclose(t["dMdxi"], ocv.dMdxi, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("ct,qo,io,vdim", [
    (fb.Quadrilateral, 2, 1, 1), (fb.Quadrilateral, 3, 2, 2), (fb.Triangle, 2, 2, 1), (fb.Triangle, 1, 1, 2),
    (fb.Hexahedron, 2, 1, 3), (fb.Hexahedron, 3, 2, 1), (fb.Tetrahedron, 1, 1, 3), (fb.Tetrahedron, 3, 2, 3),
])
def test_facetvalues_tables_match_oracle(hctx, ct, qo, io, vdim):
    # FacetQuadratureRule + facet_to_element_transformation (src/Quadrature/quadrature.jl:205-238,
    # src/FEValues/facet_integrals.jl:102-217)
    fv = fb.FacetValues(fb.FacetQuadratureRule(ct, qo), fb.Lagrange(ct, io) ** vdim, ctx=hctx)
    t = fv.tables()
    oip = O.Lagrange(SHAPE[ct], io)
    ofv = O.FacetValues(O.FacetQuadratureRule(SHAPE[ct], qo), oip ** vdim if vdim > 1 else oip)
    assert fv.nfacets == ofv.fqr.nfacets and fv.nq == ofv.w.shape[1] and fv.nbase_scalar == oip.nbase
    assert np.allclose(t["w"], ofv.w, rtol=1e-14, atol=0)
    opts = np.array([r.points for r in ofv.fqr.rules])
    assert np.allclose(t["points"], opts, rtol=1e-14, atol=1e-15)
    assert np.allclose(t["N"], ofv.N, rtol=1e-13, atol=1e-15)


def test_bad_arguments_are_reported(hctx):
    with pytest.raises(fb.FB2Error):
        fb.generate_grid(fb.Quadrilateral, (0, 2), ctx=hctx)
    with pytest.raises(fb.FB2Error):
        fb.Grid.from_arrays(fb.Quadrilateral, [(1, 2, 3, 9)], np.zeros((4, 2)), ctx=hctx)
    g = fb.generate_grid(fb.Quadrilateral, (2, 2), ctx=hctx)
    with pytest.raises(fb.FB2Error):
        fb.getfacetset(g, "nope")
    with pytest.raises(fb.FB2Error):
        fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefQuadrilateral, 3)))
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefQuadrilateral, 1) ** 2))
    ch = fb.ConstraintHandler(dh)
    with pytest.raises(fb.FB2Error):
        fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 0, [3]))
    with pytest.raises(fb.FB2Error):
        fb.ConstraintHandler.from_arrays(dh, [3, 2], [0.0, 0.0])


def test_renumber_goldens_through_cabi(hctx):
    # renumber!(dh, ch, order): the reference's literal goldens (test/test_dofs.jl:260-316) and agreement with the oracle
    from test_oracle_goldens import RENUMBER_GOLDENS

    def make():
        g = fb.generate_grid(fb.Quadrilateral, (2, 1), ctx=hctx)
        q1 = fb.Lagrange(fb.RefQuadrilateral, 1)
        dh = fb.close_(fb.add_(fb.add_(fb.DofHandler(g), "v", q1 ** 2), "s", q1))
        ch = fb.ConstraintHandler(dh)
        fb.add_(ch, fb.Dirichlet("v", fb.getfacetset(g, "left"), lambda x, t: 0, [2]))
        fb.add_(ch, fb.Dirichlet("s", fb.getfacetset(g, "left"), lambda x, t: 0))
        fb.close_(ch)
        return dh, ch

    for order, tb, c1, c2, pre in RENUMBER_GOLDENS:
        dh, ch = make()
        o = (fb.DofOrder.FieldWise if order == "fieldwise" else fb.DofOrder.ComponentWise)(tb)
        fb.renumber_(dh, ch, o)
        cd = dh.cell_dofs
        assert list(cd[0]) == c1 and list(cd[1]) == c2, order
        assert list(ch.prescribed_dofs) == pre, order
    dh, ch = make()
    cd0, pre0 = dh.cell_dofs.copy(), ch.prescribed_dofs.copy()
    perm = np.random.default_rng(1).permutation(dh.ndofs) + 1
    iperm = np.empty_like(perm)
    iperm[perm - 1] = np.arange(1, dh.ndofs + 1)
    fb.renumber_(dh, ch, perm)
    assert not np.array_equal(dh.cell_dofs, cd0)
    fb.renumber_(dh, ch, iperm)
    assert np.array_equal(dh.cell_dofs, cd0) and np.array_equal(ch.prescribed_dofs, pre0)
    with pytest.raises(fb.FB2Error):
        fb.renumber_(dh, np.ones(dh.ndofs, dtype=np.int64))            # not a permutation
    with pytest.raises(fb.FB2Error):
        fb.renumber_(dh, fb.DofOrder.FieldWise([1, 3]))                # blocks must be contiguous 1:maxblock


@pytest.mark.parametrize("ct,nel,fields", FIELDSETS)
def test_renumber_matches_oracle(hctx, ct, nel, fields):
    for order in ("fieldwise", "componentwise"):
        g, dh, og, odh = _both_dh(hctx, ct, nel, fields)
        perm = fb.renumber_(dh, (fb.DofOrder.FieldWise if order == "fieldwise" else fb.DofOrder.ComponentWise)())
        operm = O.renumber_permutation(odh, order)
        assert np.array_equal(perm, operm)
        O.renumber(odh, operm)
        assert np.array_equal(dh.cell_dofs, odh.cell_dofs)


@pytest.mark.parametrize("ct,nel,fields", FIELDSETS)
def test_update_after_renumber_matches_oracle(hctx, ct, nel, fields):
    # a ConstraintHandler closed before renumber!(dh, ch, perm) must behave like one built from the new numbering
    g, dh, og, odh = _both_dh(hctx, ct, nel, fields)
    names = sorted(og.facetsets)
    name, order, vdim = fields[0]

    def f1(x, t):
        return [np.sin(x[0] + 0.3 * k) + t + x[-1] for k in range(vdim)]

    def f2(x, t):
        return 0.25 * x[0] - x[1] + t
    both = np.concatenate([og.facetsets[names[0]][:2], og.facetsets[names[1]]])
    ch = fb.ConstraintHandler(dh)
    fb.add_(ch, fb.Dirichlet(name, fb.getfacetset(g, names[0]), f1))
    fb.add_(ch, fb.Dirichlet(name, both, f2, [1]))
    fb.close_(ch)
    perm = fb.renumber_(dh, ch, np.random.default_rng(11).permutation(dh.ndofs) + 1)
    O.renumber(odh, perm)
    och = O.ConstraintHandler(odh)
    och.add(O.Dirichlet(name, og.facetsets[names[0]], f1))
    och.add(O.Dirichlet(name, both, f2, [1]))
    och.close()
    assert np.array_equal(ch.prescribed_dofs, och.prescribed_dofs)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)
    fb.update_(ch, 0.75)
    och.update(0.75)
    assert np.array_equal(ch.inhomogeneities, och.inhomogeneities)


def test_element_assembly_needs_a_device(hctx):
    # no CPU fallback: the element-assembly entry points refuse a host-only context
    g = fb.generate_grid(fb.Quadrilateral, (2, 2), ctx=hctx)
    ip = fb.Lagrange(fb.RefQuadrilateral, 1)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.RefQuadrilateral, 2), ip, ctx=hctx)
    with pytest.raises(fb.FB2Error):
        fb.ElementAssembly(dh, cv)


def _factor_fill(cell_dofs, ndofs):
    """nnz(L + U) of a sparse LU in the given numbering (no column reordering): the quantity DofOrder.Ext{Metis} reduces"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = cell_dofs.shape[1]
    rows = np.repeat(cell_dofs[:, :, None], n, axis=2).ravel() - 1
    cols = np.repeat(cell_dofs[:, None, :], n, axis=1).ravel() - 1
    A = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(ndofs, ndofs))
    A = A + sp.identity(ndofs, format="csc") * (A.sum(axis=1).max() + 1.0)          # diagonally dominant: no pivoting
    lu = spla.splu(A, permc_spec="NATURAL", diag_pivot_thresh=0.0, options={"SymmetricMode": True})
    return lu.L.nnz + lu.U.nnz


@pytest.mark.parametrize("ct,nel,order,vdim", [(fb.Quadrilateral, (24, 24), 1, 1), (fb.Hexahedron, (8, 8, 8), 1, 1), (fb.Triangle, (10, 10), 2, 2)])
def test_renumber_metis_reduces_fill(hctx, ct, nel, order, vdim):
    # renumber!(dh, DofOrder.Ext{Metis}()) (ext/FerriteMetis.jl:29-92; the reference's test only runs it, test/test_dofs.jl:426-432)
    g = fb.generate_grid(ct, nel, ctx=hctx)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, order) ** vdim))
    cd0 = dh.cell_dofs.copy()
    scramble = np.random.default_rng(3).permutation(dh.ndofs) + 1
    fb.renumber_(dh, scramble)
    fill_scrambled = _factor_fill(dh.cell_dofs, dh.ndofs)
    perm = fb.renumber_(dh, fb.DofOrder.Metis())
    assert sorted(perm.tolist()) == list(range(1, dh.ndofs + 1))
    assert np.array_equal(dh.cell_dofs, perm[scramble[cd0 - 1] - 1])
    fill_metis = _factor_fill(dh.cell_dofs, dh.ndofs)
    fill_generated = _factor_fill(cd0, dh.ndofs)
    assert fill_metis < 0.5 * fill_scrambled, (fill_metis, fill_scrambled)
    assert fill_metis < 1.2 * fill_generated, (fill_metis, fill_generated)
    # deterministic (default METIS seed): the same graph gives the same order
    dh2 = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, order) ** vdim))
    fb.renumber_(dh2, scramble)
    assert np.array_equal(fb.renumber_(dh2, fb.DofOrder.Metis()), perm)


def test_set_inhomogeneities_arrays_in(hctx):
    g = fb.generate_grid(fb.Quadrilateral, (3, 3), ctx=hctx)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefQuadrilateral, 1)))
    ch = fb.ConstraintHandler.from_arrays(dh, [1, 5, 9], [0.0, 0.0, 0.0])
    ch.set_inhomogeneities([1.0, 2.5, -3.0])
    assert list(ch.inhomogeneities) == [1.0, 2.5, -3.0]
    with pytest.raises(fb.FB2Error):
        ch.set_inhomogeneities([1.0, 2.0])


def _dof_locations_consistent(cells, nodes, cell_dofs, oip, geo, vdim):
    """every global dof sits at one location and one component from whichever cell it is seen (test/test_dofs.jl:1117-1150)"""
    loc = {}
    for c, conn in enumerate(cells):
        x = nodes[np.asarray(conn) - 1]
        for a in range(oip.nbase):
            M, _ = geo.value_and_gradient(oip.refcoords[a])
            xd = M @ x
            for k in range(vdim):
                d = int(cell_dofs[c][a * vdim + k])
                if d in loc:
                    if not (np.allclose(loc[d][0], xd, atol=1e-12) and loc[d][1] == k):
                        return False
                else:
                    loc[d] = (xd, k)
    return True


def test_shared_face_dofs_for_all_tetrahedron_orientations(hctx):
    # test/test_dofs.jl:1099-1153 for Lagrange{RefTetrahedron, 2} (orders 3 and 4 are out of scope): 24 x 24 vertex orders
    import itertools
    nodes = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=float)
    oip, geo = O.Lagrange("tetrahedron", 2), O.Lagrange("tetrahedron", 1)
    for vdim, nshared in ((1, 6), (2, 12)):
        for t1 in itertools.permutations((1, 2, 3, 4)):
            for t2 in itertools.permutations((2, 3, 4, 5)):
                g = fb.Grid.from_arrays(fb.Tetrahedron, [t1, t2], nodes, ctx=hctx)
                dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefTetrahedron, 2) ** vdim))
                cd = dh.cell_dofs
                assert len(set(cd[0]) & set(cd[1])) == nshared and dh.ndofs == 2 * 10 * vdim - nshared
                assert _dof_locations_consistent([t1, t2], nodes, cd, oip, geo, vdim), (t1, t2)
                if vdim == 1:
                    odh = O.DofHandler(O.Grid("tetrahedron", [t1, t2], nodes)).add("u", oip).close()
                    assert np.array_equal(cd, odh.cell_dofs)


def test_shared_face_dofs_for_all_hexahedron_orientations(hctx):
    # test/test_dofs.jl:1156-1235 for Lagrange{RefHexahedron, 2}: the 24 proper rotations of both cells
    import itertools
    corner = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    rotations = []
    for cols in itertools.permutations(range(3)):
        for s in itertools.product((-1, 1), repeat=3):
            R = np.zeros((3, 3), dtype=int)
            for r in range(3):
                R[r, cols[r]] = s[r]
            if round(np.linalg.det(R)) == 1:
                rotations.append(tuple(corner.index(tuple(int(v) for v in R @ np.array(cn))) for cn in corner))
    assert len(rotations) == 24
    nodes = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1],
                      [0, 0, 2], [1, 0, 2], [1, 1, 2], [0, 1, 2]], dtype=float)
    bottom, top = (1, 2, 3, 4, 5, 6, 7, 8), (5, 6, 7, 8, 9, 10, 11, 12)
    oip, geo = O.Lagrange("hexahedron", 2), O.Lagrange("hexahedron", 1)
    for vdim, nshared in ((1, 9), (3, 27)):
        for r1 in rotations:
            for r2 in rotations:
                h1, h2 = tuple(bottom[k] for k in r1), tuple(top[k] for k in r2)
                g = fb.Grid.from_arrays(fb.Hexahedron, [h1, h2], nodes, ctx=hctx)
                dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefHexahedron, 2) ** vdim))
                cd = dh.cell_dofs
                assert len(set(cd[0]) & set(cd[1])) == nshared and dh.ndofs == 2 * 27 * vdim - nshared
                assert _dof_locations_consistent([h1, h2], nodes, cd, oip, geo, vdim), (h1, h2)
                if vdim == 1:
                    odh = O.DofHandler(O.Grid("hexahedron", [h1, h2], nodes)).add("u", oip).close()
                    assert np.array_equal(cd, odh.cell_dofs)


def test_grid_sets_by_predicate_and_by_list(hctx):
    # addfacetset! / addnodeset! / addcellset! (src/Grid/utils.jl:21-60,119-134,188-207); literals of
    # test/test_grid_dofhandler_vtk.jl:290-307 and test/test_constraints.jl:416
    g = fb.generate_grid(fb.Hexahedron, (1, 1, 1), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), ctx=hctx)
    fb.addcellset_(g, "cell_set", [1])
    fb.addnodeset_(g, "node_set", [1])
    fb.addfacetset_(g, "face_set", [(1, 1)])
    fb.addfacetset_(g, "left_face", lambda x: np.isclose(x[0], 0.0))
    assert 1 in fb.getnodeset(g, "node_set") and list(fb.getcellset(g, "cell_set")) == [1]
    assert fb.getfacetset(g, "left_face").tolist() == [[1, 5]] and fb.getfacetset(g, "face_set").tolist() == [[1, 1]]
    assert np.array_equal(fb.getfacetset(g, "left_face"), fb.getfacetset(g, "left"))
    with pytest.raises(ValueError):
        fb.addfacetset_(g, "left_face", [(1, 2)])
    g = fb.generate_grid(fb.Line, (2,), ctx=hctx)
    fb.addfacetset_(g, "center", lambda x: np.isclose(x[0], 0.0))
    assert fb.getfacetset(g, "center").tolist() == [[1, 2], [2, 1]]
    # the facet tables agree with the oracle's reference shapes, and the generated facet sets are what the predicates find
    for ct, shape, nel in ((fb.Triangle, "triangle", (3, 2)), (fb.Quadrilateral, "quadrilateral", (3, 2)),
                           (fb.Tetrahedron, "tetrahedron", (2, 2, 2)), (fb.Hexahedron, "hexahedron", (3, 2, 2))):
        from ferrite_b200 import api
        assert api._FACETS[ct] == O.REFSHAPES[shape].facets
        g = fb.generate_grid(ct, nel, ctx=hctx)
        fb.addfacetset_(g, "myleft", lambda x: np.isclose(x[0], -1.0))
        fb.addfacetset_(g, "mytop", lambda x: np.isclose(x[len(nel) - 1], 1.0))
        assert np.array_equal(fb.getfacetset(g, "myleft"), fb.getfacetset(g, "left"))
        assert np.array_equal(fb.getfacetset(g, "mytop"), fb.getfacetset(g, "top"))
        fb.addnodeset_(g, "corner", lambda x: np.allclose(x, -1.0))
        assert list(fb.getnodeset(g, "corner")) == [1]
        fb.addcellset_(g, "lower", lambda x: x[len(nel) - 1] <= 0.0)
        fb.addcellset_(g, "touch", lambda x: x[len(nel) - 1] <= -1.0, all=False)
        assert 0 < len(fb.getcellset(g, "lower")) < g.ncells and set(fb.getcellset(g, "lower")) <= set(fb.getcellset(g, "touch"))
    # a Dirichlet condition on a predicate facet set equals the one on the generated set
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(fb.RefHexahedron, 1)))
    ch1, ch2 = fb.ConstraintHandler(dh), fb.ConstraintHandler(dh)
    fb.add_(ch1, fb.Dirichlet("u", fb.getfacetset(g, "myleft"), lambda x, t: x[1]))
    fb.add_(ch2, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: x[1]))
    fb.close_(ch1)
    fb.close_(ch2)
    assert np.array_equal(ch1.prescribed_dofs, ch2.prescribed_dofs) and np.array_equal(ch1.inhomogeneities, ch2.inhomogeneities)


def test_addfacetset_name_collision_with_generated_sets_and_empty_warning():
    """src/Grid/utils.jl `_check_setname` / `_warn_emptyset`: a generated facet set name is taken; an empty set warns"""
    import warnings
    hctx = fb.Context(-1)
    g = fb.generate_grid(fb.Quadrilateral, (3, 2), ctx=hctx)
    with pytest.raises(ValueError, match="There already exists a set with the name"):
        fb.addfacetset_(g, "left", lambda x: x[0] < -0.99)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        fb.addfacetset_(g, "nothing", lambda x: x[0] > 5.0)
    assert any("no entities added" in str(x.message) for x in w)
    fb.addfacetset_(g, "mine", lambda x: x[0] < -0.99)
    assert np.array_equal(fb.getfacetset(g, "mine"), fb.getfacetset(g, "left"))

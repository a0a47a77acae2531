"""CPU tests of the partitioned (multi-GPU) path's host logic through the C ABI (host-only context).

The reference has no distributed assembly; parity for this path means: the owned columns of all ranks,
gathered, equal the serial oracle matrix.  Here the per-rank partial matrices are
produced by the oracle from each rank's OWN cells, the exchange lists come from libferrite_b200, and the
transport is emulated in-process or carried by torch.distributed/gloo with world_size 2.
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import ferrite_b200 as fb
import oracle as O

SHAPE = {fb.Quadrilateral: "quadrilateral", fb.Hexahedron: "hexahedron", fb.Tetrahedron: "tetrahedron", fb.Triangle: "triangle"}


def problem(ct, nel, order, vdim, qr, kind):
    og = O.perturb_grid(O.generate_grid(SHAPE[ct], nel), nel, (-1.0,) * len(nel), (1.0,) * len(nel), 0.2)
    oip = O.Lagrange(SHAPE[ct], order)
    oip = oip ** vdim if vdim > 1 else oip
    odh = O.DofHandler(og).add("u", oip).close()
    ocv = O.CellValues(O.QuadratureRule(SHAPE[ct], qr), oip)
    params = {"k": 1.0, "source": 1.0} if kind == "heat" else dict(zip(("lambda", "mu"), O.lame(10.0, 0.3)), b=(0.1, -0.5, 0.2)[:vdim])
    x = og.nodes[og.cells - 1]
    Ke, fe = O.ELEMENTS[kind](ocv, x, params)
    return og, odh, Ke, fe, params, ocv


def partial_matrix(odh, Ke, fe, cells0):
    """Global-size COO sum of the element matrices of the given (0-based) cells."""
    n = odh.ndofs
    cd = odh.cell_dofs[cells0] - 1
    ndpc = cd.shape[1]
    rows = np.repeat(cd[:, :, None], ndpc, axis=2).ravel()
    cols = np.repeat(cd[:, None, :], ndpc, axis=1).ravel()
    M = sp.coo_matrix((Ke[cells0].ravel(), (rows, cols)), shape=(n, n)).tocsc()
    f = np.zeros(n)
    np.add.at(f, cd.ravel(), fe[cells0].ravel())
    return M, f


def plans(hctx, ct, nel, order, vdim, nparts, dims=None):
    g = fb.generate_grid(ct, nel, ctx=hctx)
    gdh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, order) ** vdim))
    return gdh, [fb.Partition(gdh, nparts, r, dims) for r in range(nparts)]


CASES = [
    (fb.Quadrilateral, (7, 5), 1, 1, 2, "heat", 2, None),
    (fb.Quadrilateral, (6, 6), 2, 2, 3, "elasticity", 4, None),
    (fb.Hexahedron, (5, 4, 3), 1, 1, 2, "heat", 2, None),
    (fb.Hexahedron, (4, 4, 4), 1, 3, 2, "elasticity", 8, None),
    (fb.Hexahedron, (6, 3, 2), 1, 3, 2, "elasticity", 3, (3, 1, 1)),
    (fb.Hexahedron, (3, 3, 2), 2, 3, 3, "elasticity", 4, None),
    (fb.Tetrahedron, (3, 2, 2), 2, 1, 2, "heat", 4, None),
]


@pytest.mark.parametrize("ct,nel,order,vdim,qr,kind,nparts,dims", CASES)
def test_partition_plan_and_emulated_exchange(ct, nel, order, vdim, qr, kind, nparts, dims):
    hctx = fb.Context(-1)
    og, odh, Ke, fe, params, ocv = problem(ct, nel, order, vdim, qr, kind)
    gdh, parts = plans(hctx, ct, nel, order, vdim, nparts, dims)
    assert np.array_equal(gdh.cell_dofs, odh.cell_dofs)
    # own cells: a disjoint cover of all cells
    own = [p.cells_global[p.cell_is_own == 1] for p in parts]
    allown = np.concatenate(own)
    assert len(allown) == og.ncells and len(np.unique(allown)) == og.ncells
    # dof ownership: lowest rank among the cells touching the dof; consistent across ranks
    cell_owner = np.empty(og.ncells, dtype=np.int64)
    for r, c in enumerate(own):
        cell_owner[c - 1] = r
    expect_owner = np.full(odh.ndofs, nparts, dtype=np.int64)
    np.minimum.at(expect_owner, (odh.cell_dofs - 1).ravel(), np.repeat(cell_owner, odh.cell_dofs.shape[1]))
    for p in parts:
        assert np.array_equal(p.dof_owner, expect_owner[p.l2g_dof - 1])
        # local problem covers every cell touching an owned dof
        touching = np.unique(np.nonzero((expect_owner[odh.cell_dofs - 1] == p.rank).any(axis=1))[0]) + 1
        assert np.all(np.isin(touching, p.cells_global))
    assert sum(p.ndofs_owned for p in parts) == odh.ndofs
    # serial oracle
    Kser, fser = partial_matrix(odh, Ke, fe, np.arange(og.ncells))
    # per-rank partial sums from own cells, then the emulated interface exchange
    M, F = [], []
    for p in parts:
        m, f = partial_matrix(odh, Ke, fe, own[p.rank] - 1)
        M.append(m.tolil())
        F.append(f)
    msgs = {}
    for p in parts:
        for o in range(nparts):
            L = p.peer_lists(o)
            gi, gj = p.l2g_dof[L["send_rows"]] - 1, p.l2g_dof[L["send_cols"]] - 1
            if o == p.rank:
                assert len(gi) == 0 and len(L["send_f"]) == 0
                continue
            assert np.all(expect_owner[gj] == o)
            vals = np.array([M[p.rank][i, j] for i, j in zip(gi, gj)])
            fv = F[p.rank][p.l2g_dof[L["send_f"]] - 1]
            msgs[(p.rank, o)] = (gi, gj, vals, fv, p.l2g_dof[L["send_f"]] - 1)
    for p in parts:
        for s in range(nparts):
            L = p.peer_lists(s)
            if s == p.rank:
                continue
            gi, gj, vals, fv, fd = msgs[(s, p.rank)]
            # sender and receiver derived the identical list independently
            assert np.array_equal(p.l2g_dof[L["recv_rows"]] - 1, gi) and np.array_equal(p.l2g_dof[L["recv_cols"]] - 1, gj)
            assert np.array_equal(p.l2g_dof[L["recv_f"]] - 1, fd)
            for i, j, v in zip(gi, gj, vals):
                M[p.rank][i, j] += v
            F[p.rank][fd] += fv
    for p in parts:
        owned = np.nonzero(expect_owner == p.rank)[0]
        A = M[p.rank].tocsc()[:, owned]
        B = Kser[:, owned]
        d = abs(A - B)
        assert (d.max() if d.nnz else 0.0) <= 1e-12 * abs(Kser).max()
        assert np.allclose(F[p.rank][owned], fser[owned], rtol=1e-12, atol=1e-14)


def test_local_problem_reproduces_global_numbering():
    hctx = fb.Context(-1)
    gdh, parts = plans(hctx, fb.Hexahedron, (4, 3, 3), 1, 3, 4)
    gcells, gnodes, gcd = gdh.grid.cells, gdh.grid.nodes, gdh.cell_dofs
    for p in parts:
        lg, ldh = p.local_problem(hctx)
        assert lg.ncells == p.ncells_local and ldh.ndofs == p.ndofs_local
        assert np.array_equal(p.l2g_node[lg.cells - 1], gcells[p.cells_global - 1])
        assert np.array_equal(lg.nodes, gnodes[p.l2g_node - 1])
        assert np.array_equal(p.l2g_dof[ldh.cell_dofs - 1], gcd[p.cells_global - 1])
        assert np.all(np.diff(p.l2g_dof) > 0)
        # local cells: [own, interface | own, interior | halo], each ascending; own cells first
        nown = int(p.cell_is_own.sum())
        assert np.all(p.cell_is_own[:nown] == 1) and np.all(p.cell_is_own[nown:] == 0)
        assert np.all(np.diff(p.cells_global[nown:]) > 0) and len(np.unique(p.cells_global)) == len(p.cells_global)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ct, nel, order, vdim, qr, kind = fb.Hexahedron, (6, 4, 3), 1, 3, 2, "elasticity"
        og, odh, Ke, fe, params, ocv = problem(ct, nel, order, vdim, qr, kind)
        hctx = fb.Context(-1)
        g = fb.generate_grid(ct, nel, ctx=hctx)
        gdh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, order) ** vdim))
        p = fb.Partition(gdh, world, rank)
        own = p.cells_global[p.cell_is_own == 1]
        M, F = partial_matrix(odh, Ke, fe, own - 1)
        M = M.tolil()
        peer = 1 - rank
        L = p.peer_lists(peer)
        gi, gj = p.l2g_dof[L["send_rows"]] - 1, p.l2g_dof[L["send_cols"]] - 1
        send = torch.from_numpy(np.concatenate([np.array([M[i, j] for i, j in zip(gi, gj)]), F[p.l2g_dof[L["send_f"]] - 1]]))
        recv = torch.empty(len(L["recv_rows"]) + len(L["recv_f"]), dtype=torch.float64)
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)])
        for r in reqs:
            r.wait()
        recv = recv.numpy()
        ri, rj = p.l2g_dof[L["recv_rows"]] - 1, p.l2g_dof[L["recv_cols"]] - 1
        for i, j, v in zip(ri, rj, recv[:len(ri)]):
            M[i, j] += v
        F[p.l2g_dof[L["recv_f"]] - 1] += recv[len(ri):]
        Kser, fser = partial_matrix(odh, Ke, fe, np.arange(og.ncells))
        owned = p.l2g_dof[p.dof_owner == rank] - 1
        d = abs(M.tocsc()[:, owned] - Kser[:, owned])
        ok = (d.max() if d.nnz else 0.0) <= 1e-12 * abs(Kser).max() and np.allclose(F[owned], fser[owned], rtol=1e-12, atol=1e-14)
        # checksum of checksums: the owned parts of both ranks add up to the serial totals
        t = torch.tensor([M.tocsc()[:, owned].sum(), F[owned].sum(), float(len(owned))], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and abs(t[0].item() - Kser.sum()) <= 1e-9 * abs(Kser).max() and abs(t[1].item() - fser.sum()) <= 1e-10
        ok = ok and int(t[2].item()) == odh.ndofs
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_metis_partition_covers_the_grid_and_is_balanced():
    # METIS_PartMeshDual from the CUDA toolkit's libmetis_static.a, computed independently per rank (deterministic)
    hctx = fb.Context(-1)
    for ct, nel, nparts in ((fb.Hexahedron, (8, 8, 8), 4), (fb.Tetrahedron, (5, 4, 3), 3), (fb.Quadrilateral, (17, 9), 5), (fb.Triangle, (6, 6), 2)):
        g = fb.generate_grid(ct, nel, ctx=hctx)
        dh = fb.close_(fb.add_(fb.DofHandler(g), "u", fb.Lagrange(ct, 1)))
        parts = [fb.Partition(dh, nparts, r, metis=True) for r in range(nparts)]
        own = np.concatenate([p.cells_global[p.cell_is_own == 1] for p in parts])
        assert len(own) == g.ncells == len(np.unique(own))
        sizes = np.array([p.ncells_own for p in parts])
        assert sizes.min() > 0 and sizes.max() <= 1.3 * g.ncells / nparts + 2
        owned = np.concatenate([p.l2g_dof[p.dof_owner == p.rank] for p in parts])
        assert len(owned) == dh.ndofs == len(np.unique(owned))


@pytest.mark.parametrize("nel,nparts,dims,vdim,pert", [
    ((5, 4, 3), 2, None, 1, 0.2), ((6, 5, 4), 4, (2, 2, 1), 3, 0.2), ((7, 6, 5), 8, (2, 2, 2), 1, 0.0), ((9, 3, 2), 3, (3, 1, 1), 2, 0.1),
    ((3, 3, 3), 1, None, 1, 0.2), ((2, 2, 2), 8, (2, 2, 2), 1, 0.2), ((5, 7, 6), 6, (1, 2, 3), 1, 0.2), ((64, 48, 48), 2, None, 1, 0.2),
])
def test_rank_local_setup_equals_the_plan_from_the_global_problem(nel, nparts, dims, vdim, pert, monkeypatch):
    """fb2_partition_create_generated (closed forms for close!'s numbering, ownership and the node coordinates of generate_grid
    + perturb; no global grid, no global DofHandler) against fb2_partition_create on the global problem: every array of the
    plan, the exchange lists of every peer and the local grid / DofHandler must be identical."""
    monkeypatch.setenv("FB2_HOST_THREADS", "3")      # the threaded sort / translate paths
    h = fb.Context(-1)
    g = fb.generate_grid(fb.Hexahedron, nel, ctx=h)
    if pert:
        g.perturb(pert)
    ip = fb.Lagrange(fb.Hexahedron, 1) ** vdim
    gdh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))

    def plan(pt):
        d = dict(cells=pt.cells_global, own=pt.cell_is_own, l2gn=pt.l2g_node, l2gd=pt.l2g_dof, downer=pt.dof_owner,
                 info=np.array([pt.ncells_local, pt.ncells_own, pt.nnodes_local, pt.ndofs_local, pt.ndofs_owned]))
        for o in range(pt.nparts):
            d["counts%d" % o] = np.array(pt.peer_counts(o))
            for k, v in pt.peer_lists(o).items():
                d["%s%d" % (k, o)] = v
        lg, ldh = pt.local_problem(h)
        d["lcells"], d["lxyz"], d["ldofs"] = lg.cells, lg.nodes, np.asarray(ldh.cell_dofs)
        return d

    for r in range(nparts):
        a = plan(fb.Partition(gdh, nparts, r, dims))
        b = plan(fb.Partition.generated(nel, ip, nparts, r, dims, perturb=pert, host_ctx=h))
        for k in a:
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), (r, k)

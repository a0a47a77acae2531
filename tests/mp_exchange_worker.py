"""Worker of tests/test_gpu_multi.py: one rank of a real NCCL run (launched with torch.distributed.run).
Every rank assembles its local problem in EXCHANGE mode (interface cells first, NCCL send/recv of the interface
columns overlapped with the interior cells); rank 0 gathers the owned columns and compares them with the serial oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

import ferrite_b200 as fb         # noqa: E402


def main():
    out_path, kind = sys.argv[1], sys.argv[2]
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    ctx = fb.default_context(lr)
    if kind in ("heat", "heat_big"):
        # heat_big: several tiles and chunks per rank, so that the split launch of the marching kernel has both parts
        ct, nel, order, vdim, qo = fb.Hexahedron, ((12, 9, 7) if kind == "heat" else (40, 22, 26)), 1, 1, 2
        elem, oel, op = fb.HeatElement(1.5, 0.7), "heat", {"k": 1.5, "source": 0.7}
    elif kind == "elasticity_q1":                 # k_march_vec on the cell-map path
        ct, nel, order, vdim, qo = fb.Hexahedron, (22, 13, 20), 1, 3, 2
        lam, mu = 10.0 * 0.3 / (1.3 * 0.4), 10.0 / 2.6
        elem, oel, op = fb.ElasticityElement(lam=lam, mu=mu, b=(0.1, 0.2, -1.0)), "elasticity", {"lambda": lam, "mu": mu, "b": (0.1, 0.2, -1.0)}
    else:
        ct, nel, order, vdim, qo = fb.Hexahedron, (5, 4, 6), 2, 3, 3
        lam, mu = 10.0 * 0.3 / (1.3 * 0.4), 10.0 / 2.6
        elem, oel, op = fb.ElasticityElement(lam=lam, mu=mu, b=(0.1, 0.2, -1.0)), "elasticity", {"lambda": lam, "mu": mu, "b": (0.1, 0.2, -1.0)}
    gg = fb.generate_grid(ct, nel, ctx=fb.Context(-1)).perturb(0.2)
    ip = fb.Lagrange(ct, order) ** vdim
    gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(ct, qo), ip)
    part = fb.Partition(gdh, world, rank)
    g, dh = part.local_problem(ctx)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    a = fb.start_assemble(K, f)
    part.bind(a, cv)
    ids = [fb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    fb.comm_init(ctx, ids[0], world, rank)
    part._asm = a
    for _ in range(2):                       # twice: the second call must give the same result (zero fill + exchange)
        part.assemble_(elem, mode="exchange")
    ctx.synchronize()
    trip = part.owned_triplets(K, f)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(trip, gathered, dst=0)
    if rank == 0:
        import oracle as O
        shape = "hexahedron"
        og = O.perturb_grid(O.generate_grid(shape, nel), nel, (-1.0,) * 3, (1.0,) * 3, 0.2)
        oip = O.Lagrange(shape, order)
        oip = oip ** vdim if vdim > 1 else oip
        odh = O.DofHandler(og).add("u", oip).close()
        oK = O.allocate_matrix(odh)
        of = np.zeros(odh.ndofs)
        O.assemble_global(odh, O.CellValues(O.QuadratureRule(shape, qo), oip), oK, of, oel, op)
        rows = np.concatenate([t[0] for t in gathered]); cols = np.concatenate([t[1] for t in gathered])
        vals = np.concatenate([t[2] for t in gathered]); fd = np.concatenate([t[3] for t in gathered])
        fv = np.concatenate([t[4] for t in gathered])
        o = np.lexsort((rows, cols))
        ocols = np.repeat(np.arange(odh.ndofs), np.diff(oK.colptr))
        res = {
            "pattern_exact": bool(len(rows) == oK.nnz and np.array_equal(rows[o], oK.rowval) and np.array_equal(cols[o] - 1, ocols)),
            "dofs_once": bool(len(np.unique(fd)) == odh.ndofs == len(fd)),
            "nz_err": float(np.abs(vals[o] - oK.nzval).max() / np.abs(oK.nzval).max()),
            "f_err": float(np.abs(fv[np.argsort(fd)] - of).max() / max(np.abs(of).max(), 1e-300)),
            "world": world,
        }
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Breakdown of one partitioned assembly step (run under torchrun on >= 2 GPUs): milliseconds per call of
the plain local assemble (own + halo cells), own cells only, halo mode (+ mask), exchange mode (+ NCCL)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import ferrite_b200 as fb
from bench import block_dims

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
os.environ["NCCL_DEBUG"] = "WARN"
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
ctx = fb.default_context(lr)
nel = (200, 200, 200)
ip = fb.Lagrange(fb.RefHexahedron, 1)
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
elem = fb.HeatElement(1.0, 1.0)
dims = block_dims(world)
gg = fb.generate_grid(fb.Hexahedron, tuple(n * d for n, d in zip(nel, dims)), ctx=fb.Context(-1)).perturb(0.2)
gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
part = fb.Partition(gdh, world, rank, dims)
g, dh = part.local_problem(ctx)
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
a = fb.start_assemble(K, f)
part.bind(a, cv)
ids = [fb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
fb.comm_init(ctx, ids[0], world, rank)
part._asm = a
modes = {"local_all_cells": lambda: fb.assemble_(fb.start_assemble(K, f), elem, cv), "own": lambda: part.assemble_(elem, mode="own"),
         "halo": lambda: part.assemble_(elem, mode="halo"), "exchange": lambda: part.assemble_(elem, mode="exchange")}
out = {}
for name, fn in modes.items():
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=f"cuda:{lr}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = round(float(t.item()), 4)
info = part.info() if hasattr(part, "info") else None
if rank == 0:
    print(json.dumps({"world": world, "ms": out, "ncells_local": g.ncells, "info": str(info)}), file=sys.stderr)
dist.barrier(); dist.destroy_process_group()

"""One rank's local problem of a block partition, on ONE GPU: kernel time of the marching kernel through the cell map
(python scripts/part_local_time.py [world] [rank])."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ferrite_b200 as fb
from bench import block_dims
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = fb.default_context(0)
kind = os.environ.get("PL_KIND", "heat")                       # heat | elasticity
n1 = int(os.environ.get("PL_NEL", "200"))                      # cells per rank and direction
nel = (n1, n1, n1)
ip = fb.Lagrange(fb.RefHexahedron, 1) ** (3 if kind == "elasticity" else 1)
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
elem = fb.ElasticityElement(E=200e9, nu=0.3, b=(0.0, 0.0, -1.0)) if kind == "elasticity" else fb.HeatElement(1.0, 1.0)
dims = block_dims(world)
gg = fb.generate_grid(fb.Hexahedron, tuple(n * d for n, d in zip(nel, dims)), ctx=fb.Context(-1)).perturb(0.2)
gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
part = fb.Partition(gdh, world, rank, dims)
g, dh = part.local_problem(ctx)
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
a = fb.start_assemble(K, f)
part.bind(a, cv)


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


out = {"world": world, "rank": rank, "ncells_local": g.ncells, "ncells_own": part.ncells_own}
out["all_cells_accumulate"] = timed(lambda: fb.assemble_(fb.start_assemble(K, f, fillzero=False), elem, cv))
out["all_cells_step"] = timed(lambda: fb.assemble_(fb.start_assemble(K, f), elem, cv))
out["own_step"] = timed(lambda: part.assemble_(elem, mode="own"))
out["halo_step"] = timed(lambda: part.assemble_(elem, mode="halo"))
out["kernel"] = fb.last_kernel()
a._accumulate = True
out["own_accumulate"] = timed(lambda: part.assemble_(elem, mode="own"))
a._accumulate = False
a.variant = 32 if kind == "elasticity" else 30
out["own_step_percell_kernel"] = timed(lambda: part.assemble_(elem, mode="own"))
out["kernel2"] = fb.last_kernel()
print(json.dumps(out))

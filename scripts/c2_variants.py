"""C2 kernel time for a list of variants: python scripts/c2_variants.py 0 31 32 ..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ferrite_b200 as fb
ctx = fb.default_context(0)
g = fb.generate_grid(fb.Hexahedron, (200, 200, 200)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
elem = fb.HeatElement(1.0, 1.0)
out = {}
for v in [int(x) for x in sys.argv[1:]]:
    a = fb.start_assemble(K, f, fillzero=False); a.variant = v
    for _ in range(3): fb.assemble_(a, elem, cv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fb.assemble_(a, elem, cv)
    e1.record(); torch.cuda.synchronize()
    out[v] = round(e0.elapsed_time(e1) / 10, 4)
print(json.dumps(out))

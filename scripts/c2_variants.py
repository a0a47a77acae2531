"""C2 kernel time for a list of variants: python scripts/c2_variants.py 0 30 31 ...
per variant: [ms of the kernel alone (fillzero=false: everything is added), ms of a full step (zero fill + kernel)]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ferrite_b200 as fb
ctx = fb.default_context(0)
n = int(os.environ.get("C2_N", "200"))
g = fb.generate_grid(fb.Hexahedron, (n, n, n)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
elem = fb.HeatElement(1.0, 1.0)


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


def asm(v, fillzero):
    a = fb.start_assemble(K, f, fillzero=fillzero); a.variant = v
    fb.assemble_(a, elem, cv)


out = {}
for v in [int(x) for x in sys.argv[1:]]:
    out[v] = [timed(lambda: asm(v, False)), timed(lambda: asm(v, True))]
print(json.dumps(out))

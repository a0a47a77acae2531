import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ferrite_b200 as fb
v = int(sys.argv[1])
ctx = fb.default_context(0)
g = fb.generate_grid(fb.Hexahedron, (200, 200, 200)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
a = fb.start_assemble(K, f, fillzero=False); a.variant = v
for _ in range(4): fb.assemble_(a, fb.HeatElement(1.0, 1.0), cv)
torch.cuda.synchronize()

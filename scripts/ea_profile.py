"""Minimal driver for an ncu capture of k_ea_mul: python scripts/ea_profile.py [N] (Q1 hex heat, N^3 cells)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ferrite_b200 as fb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = fb.default_context(0)
g = fb.generate_grid(fb.Hexahedron, (N, N, N)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
ea = fb.ElementAssembly(dh, cv)
Kes, fes = ea.assemble(fb.HeatElement(1.0, 1.0))
x = torch.rand(dh.ndofs, dtype=torch.float64, device=Kes.device)
y = ctx.zeros(dh.ndofs)
for _ in range(3):
    ea.mul(Kes, x, out=y)
torch.cuda.synchronize()
print("ok", ea.ncells, float(y.abs().max()))

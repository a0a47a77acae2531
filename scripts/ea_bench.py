"""Element-assembly timings on one B200: python scripts/ea_bench.py [N]  (Q1 hex heat on N^3 cells, default 200 = C2).
Prints one JSON line: element matrices (fb2_ea_assemble), the matrix-free operator (fb2_ea_mul, HBM-bound: 8 n^2 bytes of Ke per
cell + dofs + x/y traffic), apply_local!, the scatter of stored matrices, and the assembled-matrix SpMV for comparison."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ferrite_b200 as fb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = fb.default_context(0)
g = fb.generate_grid(fb.Hexahedron, (N, N, N)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
K = fb.allocate_matrix(dh)
f = ctx.zeros(dh.ndofs)
elem = fb.HeatElement(1.0, 1.0)
ch = fb.ConstraintHandler(dh)
fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: 1.0))
fb.close_(ch)
ea = fb.ElementAssembly(dh, cv)
Kes, fes = ea.assemble(elem)
x = torch.rand(dh.ndofs, dtype=torch.float64, device=Kes.device)
y = ctx.zeros(dh.ndofs)
a = fb.start_assemble(K, f)
fb.assemble_(a, elem, cv)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nc, n = ea.ncells, ea.n
out = {"workload": f"heat Q1 hex {N}^3, perturbed nodes", "cells": nc, "n": n}
out["ea_assemble_ms"] = timed(lambda: ea.assemble(elem, Kes=Kes, fes=fes))
out["ea_mul_ms"] = timed(lambda: ea.mul(Kes, x, out=y))
mul_bytes = nc * (8 * n * n + 4 * n) + 8 * dh.ndofs * 3      # Ke + dofs + x read, y zero-fill + RED
out["ea_mul_GBps"] = mul_bytes / out["ea_mul_ms"] / 1e6
out["ea_mul_bytes"] = mul_bytes
out["spmv_ms"] = timed(lambda: fb.spmv(K, x, out=y))
out["assemble_ms"] = timed(lambda: fb.assemble_(fb.start_assemble(K, f), elem, cv))
out["scatter_device_ms"] = timed(lambda: fb.scatter_device_(fb.start_assemble(K, f), Kes, fes))
Kc, fc = Kes.clone(), fes.clone()
out["apply_local_ms"] = timed(lambda: ea.apply_local_(Kc, fc, ch))
out["apply_assemble_ms"] = timed(lambda: fb.apply_assemble_(fb.start_assemble(K, f), ch, elem, cv, ea=ea), reps=5, warm=2)
# consistency: operator == assembled K * x
fb.assemble_(fb.start_assemble(K, f), elem, cv)
y1 = ea.mul(Kes, x).clone()
y2 = fb.spmv(K, x).clone()
out["mul_vs_spmv_relerr"] = float((y1 - y2).norm() / y2.norm())
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = float(peaks.get("hbm_gbs") or 0) or None
except Exception:
    hbm = None
out["hbm_peak_GBps"] = hbm
if hbm:
    out["ea_mul_frac_of_hbm"] = out["ea_mul_GBps"] / hbm
print(json.dumps(out))

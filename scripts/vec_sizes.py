"""k_march_vec / k_march_hex (VS_KIND=heat) on generated grids of several sizes (python scripts/vec_sizes.py 128x128x128 129x128x128 ...): ms per launch"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ferrite_b200 as fb
ctx = fb.default_context(0)
kind = os.environ.get("VS_KIND", "elasticity")
ip = fb.Lagrange(fb.RefHexahedron, 1) ** (3 if kind == "elasticity" else 1)
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
elem = fb.ElasticityElement(E=200e9, nu=0.3, b=(0.0, 0.0, -1.0)) if kind == "elasticity" else fb.HeatElement(1.0, 1.0)
for spec in sys.argv[1:]:
    nel = tuple(int(v) for v in spec.split("x"))
    g = fb.generate_grid(fb.Hexahedron, nel).perturb(0.2)
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
    def run():
        a = fb.start_assemble(K, f, fillzero=False)
        fb.assemble_(a, elem, cv)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"nel": nel, "ms": round(ms, 4), "ns_per_cell": round(ms * 1e6 / g.ncells, 4), "kernel": fb.last_kernel()}), flush=True)
    del K, f, dh, g
    torch.cuda.empty_cache()

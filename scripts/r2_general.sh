timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "general_stiffness or marching_tile_kernel_elasticity or incomplete" 2>&1 | tail -3
python - <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
import ferrite_b200 as fb, oracle as O
ctx = fb.default_context(0)
ip = fb.Lagrange(fb.RefHexahedron, 1) ** 3
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
g = fb.generate_grid(fb.Hexahedron, (128, 128, 128)).perturb(0.2)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
K = fb.allocate_matrix(dh); f = ctx.zeros(dh.ndofs)
lam, mu = O.lame(200e9, 0.3)
el = fb.GeneralElasticityElement(O.isotropic_stiffness(lam, mu, 3), (0.0, 0.0, -1.0))
for variant in (0, 1):
    def run():
        a = fb.start_assemble(K, f); a.variant = variant
        fb.assemble_(a, el, cv)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"general_C_128^3_step_ms": round(e0.elapsed_time(e1) / 10, 3), "variant": variant, "kernel": fb.last_kernel()}))
PY

"""Time the host-buffer entry points on C2: plain, streamed without / with the coordinate upload, for several slab counts."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ferrite_b200 as fb
ctx = fb.default_context(0)
g = fb.generate_grid(fb.Hexahedron, (200, 200, 200)).perturb(0.2)
ip = fb.Lagrange(fb.RefHexahedron, 1)
dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
cv = fb.CellValues(fb.QuadratureRule(fb.RefHexahedron, 2), ip)
K = fb.allocate_matrix(dh)
elem = fb.HeatElement(1.0, 1.0)
xyz = torch.from_numpy(g.nodes).pin_memory(); nz = torch.empty(K.nnz, dtype=torch.float64).pin_memory(); f = torch.empty(dh.ndofs, dtype=torch.float64).pin_memory()
a = fb.start_assemble(K, None)
def t(fn, n=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
out = {}
def plain():
    g.upload_coordinates_async(xyz); fb.assemble_host(fb.start_assemble(K, None), elem, cv, nz.numpy(), f.numpy())
out["plain_upload+assemble_host"] = t(plain)
out["assemble_host_only"] = t(lambda: fb.assemble_host(fb.start_assemble(K, None), elem, cv, nz.numpy(), f.numpy()))
for ns in (1, 2, 4, 8, 16):
    os.environ["FB2_HOST_SLABS"] = str(ns)
    out[f"streamed_noxyz_{ns}"] = t(lambda: fb.assemble_host_streamed(fb.start_assemble(K, None), elem, cv, nz.numpy(), f.numpy()))
    out[f"streamed_xyz_{ns}"] = t(lambda: fb.assemble_host_streamed(fb.start_assemble(K, None), elem, cv, nz.numpy(), f.numpy(), xyz=xyz.numpy()))
# raw copies
d = torch.empty(K.nnz, dtype=torch.float64, device="cuda")
out["raw_d2h_nzval"] = t(lambda: nz.copy_(d, non_blocking=True))
print(json.dumps({k: round(v, 2) for k, v in out.items()}))

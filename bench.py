#!/usr/bin/env python
"""bench.py -- cells assembled/s (K+f, FP64, 3-D hex) on N B200s, one JSON line on stdout.

Workload at N=1: BASELINE.json configs[1] = heat equation, scalar Q1 on generate_grid(Hexahedron,(200,200,200)),
nodes perturbed deterministically so that cells are not congruent, K and f assembled
in FP64 (zero-fill + gather + geometry + element + scatter).  A "step" is one full assembly.

  value        device-resident throughput: cells / s with grid, dofs, pattern and map already in HBM
  e2e          same metric through the host-buffer C-ABI path: per step the node coordinates are copied
               host(pinned)->device and nzval + f are copied device->host(pinned)
  roofline     dominant kernel (fused cell kernel) against the measured HBM peak, algorithmic bytes per cell
               from SURVEY.md section 8d / DESIGN.md (B_min = 314 B/cell for this config); the FP64 view
               (F_min = 4.3 kFLOP/cell against an FMA microbenchmark on the same device) is reported beside it
  cpu_baseline the oracle's C restatement of the reference loop (oracle/cpu_assemble.c) on the host cores, on a
               bounded sample of the same workload
`--impl reference` times that CPU restatement alone (the reference is Julia; no Julia toolchain exists in the
image or on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (celltype, nel, order, vdim, qr_order, element, B_min bytes/cell, F_min flop/cell)
    "c1": dict(cell="quad", nel=(100, 100), order=1, vdim=1, qr=2, element="heat", bmin=126.0, fmin=0.45e3,
               label="heat Q1 quad 100^2 (BASELINE.json configs[0], the reference's CPU tutorial path; 10 000 cells: "
                     "launch-latency bound, report the time, not the roofline)"),
    "c2": dict(cell="hex", nel=(200, 200, 200), order=1, vdim=1, qr=2, element="heat", bmin=314.0, fmin=4.3e3,
               label="heat Q1 hex 200^3 (BASELINE.json configs[1]), perturbed nodes"),
    "c2s": dict(cell="hex", nel=(64, 64, 64), order=1, vdim=1, qr=2, element="heat", bmin=314.0, fmin=4.3e3,
                label="heat Q1 hex 64^3 (reduced size, debugging only)"),
    "c5": dict(cell="hex", nel=(128, 128, 128), order=1, vdim=3, qr=2, element="elasticity", bmin=2060.0, fmin=15.5e3,
               label="linear elasticity Q1^3 hex 128^3 (per-GPU block of BASELINE.json configs[4])"),
    "c5full": dict(cell="hex", nel=(160, 160, 160), order=1, vdim=3, qr=2, element="elasticity", bmin=2060.0, fmin=15.5e3,
                   label="linear elasticity Q1^3 hex, 160^3 cells per GPU (BASELINE.json configs[4]: 320^3 cells on 8 GPUs)"),
    "c3": dict(cell="hex", nel=(48, 48, 48), order=2, vdim=3, qr=3, element="elasticity", bmin=37.6e3, fmin=0.48e6,
               label="linear elasticity Q2^3 hex 48^3 (BASELINE.json configs[2] at 1/8 size)"),
    # bytes: 349 stored entries per cell (nnz / ncells of this mesh) written once = 2.79 kB, + dofs 120 + conn 16 + u, f, x 80;
    # flops per quadrature point (11 of them): grad u 90 FMA, constitutive law + dP/dF ~400, residual 30 x 3, tangent
    # 30 x 27 + 900 x 3 (exploiting that grad(delta u_i) has one non-zero row; the reference's dense tensor contractions
    # perform 10.8 kFMA per point = 0.24 MFLOP per cell): 4.1 kFMA x 11 x 2 = 0.09 MFLOP
    "c4": dict(cell="tet", nel=(48, 48, 48), order=2, vdim=3, qr=4, element="neohooke", bmin=3.0e3, fmin=0.09e6,
               label="Neo-Hooke tangent + residual, P2^3 tetrahedra, generate_grid(Tetrahedron, 48^3) = 663552 cells "
                     "(BASELINE.json configs[3])"),
    "c3full": dict(cell="hex", nel=(96, 96, 96), order=2, vdim=3, qr=3, element="elasticity", bmin=37.6e3, fmin=0.48e6,
                   label="linear elasticity Q2^3 hex 96^3 (BASELINE.json configs[2], full size: nnz = 4.09e9 > 2^32)"),
}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end measurement are first-touched on the GPU's NUMA node (torchrun does not bind ranks)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:            # binding is an optimisation only
        print(f"[bench] NUMA binding skipped: {e}", file=sys.stderr)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def SAMPLE(cfg):
    """what the CPU arm assembles per step: the FULL configuration for first-order hexahedra / quadrilaterals (the C
    restatement sets 200^3 up in seconds, oracle/cpu_setup.c), a bounded sample (cubes per direction) otherwise"""
    if cfg["element"] == "neohooke":
        return (24, 24, 24)
    if len(cfg["nel"]) == 2:
        return tuple(cfg["nel"])
    if cfg["order"] == 1 and cfg.get("cell", "hex") == "hex":
        return tuple(cfg["nel"])
    return (16, 16, 16)


def oracle_problem(cfg, nel):
    import oracle as O
    shape = {"tet": "tetrahedron", "quad": "quadrilateral"}.get(cfg.get("cell"), "hexahedron")
    dim = len(nel)
    if shape == "hexahedron" and cfg["order"] == 1:
        # first-order hexahedra: grid, numbering and pattern from the C restatement (identical arrays, seconds at 200^3)
        from oracle import cport
        og, dh, K = cport.hex_q1_problem(nel, cfg["vdim"])
        ip = O.Lagrange(shape, 1)
        ip = ip ** cfg["vdim"] if cfg["vdim"] > 1 else ip
        dh._K = K
        return og, dh, O.CellValues(O.QuadratureRule(shape, cfg["qr"]), ip)
    og = O.perturb_grid(O.generate_grid(shape, nel), nel, (-1,) * dim, (1,) * dim, 0.2)
    ip = O.Lagrange(shape, cfg["order"])
    ip = ip ** cfg["vdim"] if cfg["vdim"] > 1 else ip
    dh = O.DofHandler(og).add("u", ip).close()
    cv = O.CellValues(O.QuadratureRule(shape, cfg["qr"]), ip)
    return og, dh, cv


def cpu_assembler(cfg, dh, cv, K, f, colors=None):
    """(callable doing one CPU assembly, threads used, description): the C/OpenMP restatement of the reference loop;
    colors = cport.structured_coloring(...) selects the coloured scheme of the threaded how-to instead of the atomic one."""
    import numpy as np
    import oracle as O
    from oracle import cport
    what = ("C restatement of the reference's threaded atomic loop" if colors is None else
            f"C restatement of the reference's threaded coloured loop, {colors[0]} colours")
    if cfg["element"] == "neohooke":
        E, nu = 10.0, 0.3
        params = {"mu": E / (2 * (1 + nu)), "lambda": E * nu / ((1 + nu) * (1 - 2 * nu)), "b": (0.0, -0.5, 0.0)}
        u = 1e-3 * np.sin(0.37 * np.arange(dh.ndofs, dtype=np.float64))
        nthreads = os.cpu_count() or 1
        return (lambda: cport.assemble(dh, cv, K, f, "neohooke", params, nthreads=nthreads, u=u, colors=colors)), nthreads, what
    if cfg["element"] == "heat":
        params = {"k": 1.0, "source": 1.0}
    else:
        lam, mu = O.lame(200e9, 0.3)
        params = {"lambda": lam, "mu": mu, "b": (0.0, 0.0, -1.0)}
    nthreads = os.cpu_count() or 1
    return (lambda: cport.assemble(dh, cv, K, f, cfg["element"], params, nthreads=nthreads, colors=colors)), nthreads, what


def cpu_variants(cfg, og, dh, cv, K, f, sample_nel):
    """The two threading schemes of docs/src/literate-howto/threaded_assembly.jl (atomic adds; colours): [(run, threads, what)]"""
    from oracle import cport
    out = [cpu_assembler(cfg, dh, cv, K, f)]
    try:
        out.append(cpu_assembler(cfg, dh, cv, K, f, colors=cport.structured_coloring(og, sample_nel)))
    except Exception:          # no structured colouring for this grid: the atomic scheme alone
        pass
    return out


def cpu_baseline(cfg, sample_nel, reps=3, threads=0):
    """Time the C restatement of the reference loop on a bounded sample of the workload (all host cores)."""
    import numpy as np
    import oracle as O
    from oracle import cport
    og, dh, cv = oracle_problem(cfg, sample_nel)
    K = getattr(dh, "_K", None) or O.allocate_matrix(dh)
    f = np.zeros(dh.ndofs)
    best, nthreads, what, others = float("inf"), 1, "", []
    for run, nt, w in cpu_variants(cfg, og, dh, cv, K, f, sample_nel):     # both schemes of the how-to, the faster one is reported
        run()   # warm-up
        b = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            run()
            b = min(b, time.perf_counter() - t0)
        others.append(f"{w}: {og.ncells / b:.3g} cells/s")
        if b < best:
            best, nthreads, what = b, nt, w
    return {"value": og.ncells / best, "unit": "cells/s", "cores": nthreads, "kind": "port",
            "sample": f"{'x'.join(map(str, sample_nel))} cubes of the same workload"
                      f"{' = the full configuration' if tuple(sample_nel) == tuple(cfg['nel']) else ''} ({what}, best of {reps}; measured: "
                      f"{'; '.join(others)}); the reference is Julia and cannot run here"}


def run_reference(args, cfg):
    """--impl reference: the reference's CPU algorithm (oracle C port) on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = SAMPLE(cfg)
    import numpy as np
    import oracle as O
    og, dh, cv = oracle_problem(cfg, sample)
    K = getattr(dh, "_K", None) or O.allocate_matrix(dh)
    f = np.zeros(dh.ndofs)
    # the faster of the how-to's two threading schemes (one trial run each) is the one timed
    trials = []
    for cand in cpu_variants(cfg, og, dh, cv, K, f, sample):
        cand[0]()
        t0 = time.perf_counter()
        cand[0]()
        trials.append((time.perf_counter() - t0, cand))
    t_best, (run, nthreads, what) = min(trials, key=lambda t: t[0])
    # the whole run must end within a few minutes whatever the host: if the full configuration would take longer, assemble a
    # slab of it (fewer layers of the same cross-section) per step instead and say so
    budget_s = float(os.environ.get("FB2_REF_BUDGET_S", "150"))
    projected = t_best * (args.steps + args.warmup)
    if projected > budget_s and len(sample) == 3 and sample[2] > 8:
        nz_s = max(8, int(sample[2] * budget_s / projected))
        sample = (sample[0], sample[1], nz_s)
        og, dh, cv = oracle_problem(cfg, sample)
        K = getattr(dh, "_K", None) or O.allocate_matrix(dh)
        f = np.zeros(dh.ndofs)
        cands = cpu_variants(cfg, og, dh, cv, K, f, sample)
        run, nthreads, what = next((c for c in cands if c[2] == what), cands[0])
        run()
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = og.ncells * args.steps / dt
    sample_txt = (f"{'x'.join(map(str, sample))} cells per step of the same workload"
                  + (" = the full configuration" if tuple(sample) == tuple(cfg["nel"]) else ""))
    line = {
        "impl": "reference", "metric": "cells assembled/sec (K+f, FP64, 3D hex)", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": nthreads, "kind": "port", "sample": sample_txt +
                         f"; {what} (the reference is Julia, no toolchain here)"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def nccl_vs_oracle_check(fb, ctx, dist, world, rank, per_rank=24):
    """N > 1, before the timing: a small heat problem (per_rank^3 cells per GPU) assembled through the real NCCL exchange
    path, owned columns gathered on rank 0 and compared entry by entry with the serial oracle (the C restatement of the
    reference loop; the oracle is the checker here, outside every timed region).  Returns {nccl_vs_oracle_nz_err, _f_err,
    nccl_vs_oracle_pattern_exact} on rank 0, {} elsewhere."""
    import numpy as np
    dims = block_dims(world)
    nel = tuple(per_rank * d for d in dims)
    hctx = fb.Context(-1)
    gg = fb.generate_grid(fb.Hexahedron, nel, ctx=hctx).perturb(0.2)
    ip = fb.Lagrange(fb.Hexahedron, 1)
    gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
    cv = fb.CellValues(fb.QuadratureRule(fb.Hexahedron, 2), ip)
    part = fb.Partition(gdh, world, rank, dims)
    g, dh = part.local_problem(ctx)
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    part.bind(fb.start_assemble(K, f), cv)
    elem = fb.HeatElement(1.5, 0.7)
    for _ in range(2):                      # twice: zero fill + exchange must be repeatable
        part.assemble_(elem, mode="exchange")
    ctx.synchronize()
    trip = part.owned_triplets(K, f)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(trip, gathered, dst=0)
    out = {}
    if rank == 0:
        import oracle as O
        from oracle import cport
        og = O.perturb_grid(O.generate_grid("hexahedron", nel), nel, (-1.0,) * 3, (1.0,) * 3, 0.2)
        oip = O.Lagrange("hexahedron", 1)
        odh = O.DofHandler(og).add("u", oip).close()
        oK = O.allocate_matrix(odh)
        of = np.zeros(odh.ndofs)
        cport.assemble(odh, O.CellValues(O.QuadratureRule("hexahedron", 2), oip), oK, of, "heat", {"k": 1.5, "source": 0.7}, nthreads=1)
        rows = np.concatenate([t[0] for t in gathered])
        cols = np.concatenate([t[1] for t in gathered])
        vals = np.concatenate([t[2] for t in gathered])
        fd = np.concatenate([t[3] for t in gathered])
        fv = np.concatenate([t[4] for t in gathered])
        o = np.lexsort((rows, cols))
        ocols = np.repeat(np.arange(odh.ndofs), np.diff(oK.colptr))
        exact = bool(len(rows) == oK.nnz and np.array_equal(rows[o], oK.rowval) and np.array_equal(cols[o] - 1, ocols)
                     and len(np.unique(fd)) == odh.ndofs == len(fd))
        out = {"nccl_vs_oracle_pattern_exact": exact,
               "nccl_vs_oracle_nz_err": float(np.abs(vals[o] - oK.nzval).max() / np.abs(oK.nzval).max()) if exact else None,
               "nccl_vs_oracle_f_err": float(np.abs(fv[np.argsort(fd)] - of).max() / np.abs(of).max()) if exact else None,
               "nccl_vs_oracle_problem": f"heat Q1 hex {'x'.join(map(str, nel))}, {world} ranks, exchange mode"}
    dist.barrier()
    del part, K, f, g, dh
    return out


def config4_strong(fb, ctx, dist, torch, world, rank, dev, steps, nel_total=(320, 320, 320)):
    """BASELINE.json configs[4]: Q1^3 linear elasticity on 320^3 cells in total, partitioned over the N ranks (strong
    scaling: the global problem is fixed), NCCL exchange of the interface columns.  Returns the record on rank 0."""
    import time as _t
    rec = {"workload": f"linear elasticity Q1^3 hex {'x'.join(map(str, nel_total))} in total (BASELINE.json configs[4]), "
                       f"block partition over {world} ranks", "scaling": "strong"}
    try:
        t0 = _t.perf_counter()
        dims = block_dims(world)
        hctx = fb.Context(-1)
        ip = fb.Lagrange(fb.Hexahedron, 1) ** 3
        cv = fb.CellValues(fb.QuadratureRule(fb.Hexahedron, 2), ip)
        if not os.environ.get("FB2_BENCH_GLOBAL_SETUP"):
            part = fb.Partition.generated(nel_total, ip, world, rank, dims, perturb=0.2, host_ctx=hctx)   # rank-local set-up
            ncells = int(nel_total[0] * nel_total[1] * nel_total[2])
        else:
            gg = fb.generate_grid(fb.Hexahedron, nel_total, ctx=hctx).perturb(0.2)
            gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
            part = fb.Partition(gdh, world, rank, dims)
            ncells = gg.ncells
            del gdh, gg
        g, dh = part.local_problem(ctx)
        K = fb.allocate_matrix(dh)
        f = ctx.zeros(dh.ndofs)
        part.bind(fb.start_assemble(K, f), cv)
        elem = fb.ElasticityElement(E=200e9, nu=0.3, b=(0.0, 0.0, -1.0))
        for _ in range(2):
            part.assemble_(elem, mode="exchange")
        torch.cuda.synchronize()
        rec["setup_s"] = round(_t.perf_counter() - t0, 1)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            part.assemble_(elem, mode="exchange")
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fz = torch.tensor([float(f[2::3].sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(fz, op=dist.ReduceOp.SUM)
        rec.update({"ms_per_step": float(t.item()), "cells_per_s": ncells / (float(t.item()) * 1e-3), "cells": ncells,
                    "nnz_rank0": K.nnz, "steps": steps, "sum_fz_plus_volume": abs(float(fz.item()) + 8.0)})
        del part, K, f
    except Exception as exc:                 # an extra record must never cost the headline line
        rec["error"] = str(exc)[:300]
        try:
            dist.barrier()
        except Exception:
            pass
    torch.cuda.empty_cache()
    return rec


def block_dims(n):
    """px, py, pz with px*py*pz == n, as cubic as possible (2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2)."""
    dims = [1, 1, 1]
    p = 2
    while n > 1:
        if n % p:
            p += 1
            continue
        dims[dims.index(min(dims))] *= p
        n //= p
    return tuple(sorted(dims, reverse=True))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "colored"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--dist-mode", default="exchange", choices=["exchange", "halo"],
                    help="N>1: own cells + NCCL interface exchange, or own+halo cells without communication")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="N>1: skip the extra strong-scaling record of BASELINE.json configs[4]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    import numpy as np
    import torch
    import ferrite_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL's version banner) go to stderr meanwhile
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # NCCL's communicator / transport lines go to stderr (fd 1 is redirected above), so the driver can see them
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    bind_to_gpu_numa_node(local_rank)
    ctx = fb.default_context(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    nccl_check = {}
    if world > 1:
        # the library's own NCCL communicator (interface exchange); created once, used by every partitioned problem below
        ids = [fb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        fb.comm_init(ctx, ids[0], world, rank)
        nccl_check = nccl_vs_oracle_check(fb, ctx, dist, world, rank)

    # ---- set-up (not timed): grid, dofs, pattern, map all resident in HBM ------------------------------
    t_setup = time.perf_counter()
    nel = cfg["nel"]
    ct = {"tet": fb.Tetrahedron, "quad": fb.Quadrilateral}.get(cfg["cell"], fb.Hexahedron)
    ip = fb.Lagrange(ct, cfg["order"]) ** cfg["vdim"]
    cv = fb.CellValues(fb.QuadratureRule(ct, cfg["qr"]), ip)
    if cfg["element"] == "heat":
        elem = fb.HeatElement(1.0, 1.0)
    elif cfg["element"] == "elasticity":
        elem = fb.ElasticityElement(E=200e9, nu=0.3, b=(0.0, 0.0, -1.0))
    else:
        elem = fb.NeoHookeElement(E=10.0, nu=0.3, b=(0.0, -0.5, 0.0))      # hyperelasticity.jl:334-338
    u_state = None
    part = None
    if world == 1:
        g = fb.generate_grid(ct, nel).perturb(0.2)
        dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
        ncells_total, volume = g.ncells, 2.0 ** len(nel)
    else:
        # weak scaling: every GPU owns one block of `nel` cells of a px*py*pz times larger box; every rank derives
        # the global numbering (host), its own + halo cells and the interface exchange lists
        dims = block_dims(world)
        gnel = tuple(n * d for n, d in zip(nel, dims))
        hctx = fb.Context(-1)
        gleft, gright = tuple(-float(d) for d in dims), tuple(float(d) for d in dims)
        if cfg["cell"] == "hex" and cfg["order"] == 1 and not os.environ.get("FB2_BENCH_GLOBAL_SETUP"):
            # rank-local set-up: the same plan as Partition(close!(DofHandler(generate_grid(...).perturb(0.2))), world, rank, dims)
            # from closed forms, without the global grid / DofHandler (tests/test_partition_host.py compares the two)
            part = fb.Partition.generated(gnel, ip, world, rank, dims, left=gleft, right=gright, perturb=0.2, host_ctx=hctx)
            ncells_total = int(gnel[0] * gnel[1] * gnel[2])
        else:
            gg = fb.generate_grid(fb.Hexahedron, gnel, gleft, gright, ctx=hctx).perturb(0.2)
            gdh = fb.close_(fb.add_(fb.DofHandler(gg), "u", ip))
            part = fb.Partition(gdh, world, rank, dims)
            ncells_total = gg.ncells
        g, dh = part.local_problem(ctx)
        volume = 8.0 * world
    K = fb.allocate_matrix(dh)
    f = ctx.zeros(dh.ndofs)
    if cfg["element"] == "neohooke":
        # a deterministic small displacement state (|grad u| ~ 0.03, det F > 0); the work per cell does not depend on it
        u_state = (1e-3 * torch.sin(0.37 * torch.arange(dh.ndofs, dtype=torch.float64))).to(dev)
    def new_assembler(f_vec=f, fillzero=True):
        """start_assemble(K, f): zero fill pending for the next assemble call (the native assembler is cached on K)"""
        asm = fb.start_assemble(K, f_vec, fillzero=fillzero, scatter=args.scatter)
        asm.variant = args.variant
        return asm

    a = new_assembler()
    if part is not None:
        part.bind(a, cv)

    def step(asm=None):
        """one full assembly: start_assemble (zero fill) + the cell loop [+ interface exchange]"""
        if part is None:
            fb.assemble_(new_assembler(), elem, cv, u=u_state)
        else:
            part.assemble_(elem, mode=args.dist_mode)

    step(a)
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step(a)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(a)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    ctx.synchronize()
    value = ncells_total * args.steps / (ms * 1e-3)
    # size-independent checks on the full-size result (Laplace: sum(f) = |Omega| * source, sum(K) = 0); with N > 1
    # every rank holds its owned columns / dofs only, so the sums over ranks must give the global totals
    checks = {}
    if cfg["element"] == "heat":
        checks["sum_f_minus_volume"] = abs(sum_over_ranks(float(f.sum())) - volume)
        checks["sum_K"] = abs(sum_over_ranks(float(K.nzval.sum()))) / max(max_over_ranks(float(K.nzval.abs().max())), 1e-300)
    if cfg["element"] == "elasticity" and world == 1:
        # rigid-body translation is in the null space of K (K symmetric: K't = Kt), and sum(f_z) = -|Omega| for b = (0,0,-1)
        tvec = torch.zeros(K.n, dtype=torch.float64, device=dev)
        tvec[0::3] = 1.0
        yv = fb.spmv(K, tvec, transpose=True)
        checks["K_times_translation_rel"] = float(yv.abs().max()) / float(K.nzval.abs().max())
        checks["sum_fz_plus_volume"] = abs(float(f[2::3].sum()) + volume)
        del tvec, yv
    apply_info = None
    if cfg["element"] == "elasticity" and world == 1:
        # BASELINE.json configs[2] "... with Dirichlet apply!": u = 0 on "left", a smooth non-zero field on "right"
        ch = fb.ConstraintHandler(dh)
        fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "left"), lambda x, t: [0.0, 0.0, 0.0], [1, 2, 3]))
        fb.add_(ch, fb.Dirichlet("u", fb.getfacetset(g, "right"), lambda x, t: [0.0, 0.0, 0.01 * x[1]], [1, 2, 3]))
        fb.close_(ch)
        fb.update_(ch, 0.0)
        ctx.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        fb.apply_(K, f, ch)
        a1.record()
        torch.cuda.synchronize()
        apply_info = {"ms": a0.elapsed_time(a1), "prescribed_dofs": int(len(ch.prescribed_dofs))}
        step(a)            # restore the unconstrained K, f for the remaining measurements
        ctx.synchronize()
    if part is not None:
        # the exchange path and the communication-free halo path must give the same owned columns
        ref_nz, ref_f = K.nzval.clone(), f.clone()
        part.assemble_(elem, mode="halo" if args.dist_mode == "exchange" else "exchange")
        ctx.synchronize()
        scale = max_over_ranks(float(ref_nz.abs().max()))
        checks["exchange_vs_halo_maxdiff_rel"] = max_over_ranks(float((K.nzval - ref_nz).abs().max())) / scale
        checks["exchange_vs_halo_f_maxdiff"] = max_over_ranks(float((f - ref_f).abs().max()))
        del ref_nz, ref_f
        step(a)
        ctx.synchronize()

    # ---- dominant kernel alone (no zero fill, no exchange) for the roofline -----------------------------------
    a_nz = new_assembler(fillzero=False)
    kreps = max(3, min(args.steps, 10))
    fb.assemble_(a_nz, elem, cv, u=u_state)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(kreps):
        fb.assemble_(a_nz, elem, cv, u=u_state)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / kreps
    step(a)       # restore K, f (the accumulating launches above added on top)
    ctx.synchronize()
    fsum_dev = float(f.sum())

    # ---- end-to-end through host buffers ------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e and K.nnz * 8 <= 8e9:   # larger results: the pinned host buffer alone would exceed 8 GB
        xyz_host = torch.from_numpy(g.nodes).pin_memory()
        nz_host = torch.empty(K.nnz, dtype=torch.float64).pin_memory()
        f_host = torch.empty(dh.ndofs, dtype=torch.float64).pin_memory()
        nz_np, f_np = nz_host.numpy(), f_host.numpy()
        esteps = max(2, min(args.steps, 5))

        xyz_np = xyz_host.numpy()
        u_np = u_state.cpu().numpy() if u_state is not None else None

        def e2e_step():
            if part is None:
                # H2D of this step's coordinates, assembly and D2H of nzval + f, pipelined over 8 slabs of cells
                # (synchronises before returning)
                fb.assemble_host_streamed(new_assembler(None), elem, cv, nz_np, f_np, xyz=xyz_np, u=u_np)
            else:
                g.upload_coordinates_async(xyz_host)          # H2D: this step's input
                step()
                nz_host.copy_(K.nzval, non_blocking=True)     # D2H of this rank's owned columns / dofs
                f_host.copy_(f, non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": ncells_total * esteps / dt, "unit": "cells/s",
               "h2d_bytes_per_step": int(sum_over_ranks(float(g.nnodes * g.sdim * 8))),
               "d2h_bytes_per_step": int(sum_over_ranks(float((K.nnz + dh.ndofs) * 8))), "steps": esteps,
               "checksum_ok": bool(abs(float(f_host.sum()) - fsum_dev) <= 1e-9 * max(1.0, abs(fsum_dev)))}
        if world > 1:
            # the ceiling of this path on this box: the same D2H copy alone, all ranks at once (the step is D2H-bound: the
            # copy of nzval is >= 85 % of its bytes), worst rank
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                nz_host.copy_(K.nzval, non_blocking=True)
                torch.cuda.synchronize()
            dt_raw = max_over_ranks(time.perf_counter() - t0) / 2
            e2e["d2h_alone_GBps_per_gpu"] = round(K.nnz * 8 / dt_raw / 1e9, 1)
            e2e["d2h_in_step_GBps_per_gpu"] = round((K.nnz + dh.ndofs) * 8 * esteps / dt / 1e9, 1)
            e2e["cpu_affinity"] = len(os.sched_getaffinity(0))

    # ---- the consumer of K: y = K x on the assembled matrix, an HBM-bound gather ------------------------
    spmv = None
    if world == 1:
        xv = torch.ones(K.n, dtype=torch.float64, device=dev)
        yv = torch.empty_like(xv)
        fb.spmv(K, xv, transpose=True, out=yv)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(10):
            fb.spmv(K, xv, transpose=True, out=yv)
        s1.record()
        torch.cuda.synchronize()
        spmv_ms = s0.elapsed_time(s1) / 10
        spmv_bytes = K.nnz * 12 + K.n * (8 + 8 + 8)       # nzval + rowval, colptr, x once, y
        spmv = {"ms": spmv_ms, "GB/s": spmv_bytes / spmv_ms / 1e6, "bytes": spmv_bytes,
                "rowsum_check": float(yv.abs().max()) / float(K.nzval.abs().max()) if cfg["element"] == "heat" else None}
        del xv, yv

    # ---- element assembly: stored element matrices and the matrix-free operator y = sum_e P' Ke P x ------
    ea_info = None
    if world == 1 and cfg["element"] == "heat" and g.ncells * dh.ndofs_per_cell ** 2 * 8 <= 8 << 30:
        try:
            ea = fb.ElementAssembly(dh, cv)
            Kes, fes = ea.assemble(elem)
            xv = torch.ones(dh.ndofs, dtype=torch.float64, device=dev)
            yv = torch.empty_like(xv)

            def _ms(fn, reps=10):
                fn()
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(reps):
                    fn()
                t1.record()
                torch.cuda.synchronize()
                return t0.elapsed_time(t1) / reps
            n_loc = dh.ndofs_per_cell
            mul_ms = _ms(lambda: ea.mul(Kes, xv, out=yv))
            mul_bytes = g.ncells * (8 * n_loc * n_loc + 4 * n_loc) + 3 * 8 * dh.ndofs    # Ke + dofs, x read, y zero-fill + RED
            ea_info = {"element_matrices_ms": _ms(lambda: ea.assemble(elem, Kes=Kes, fes=fes)), "mul_ms": mul_ms,
                       "mul_GB/s": mul_bytes / mul_ms / 1e6, "mul_bytes": mul_bytes,
                       "rowsum_check": float(yv.abs().max()) / float(Kes.abs().max())}
            del ea, Kes, fes, xv, yv
        except Exception as exc:          # an extra measurement must never cost the headline line
            ea_info = {"error": str(exc)[:200]}

    ndofs0, nnz0, gcells0 = dh.ndofs, K.nnz, g.ncells
    config4 = None
    if world > 1 and args.config == "c2" and not args.no_config4:
        # free the headline problem first: configs[4] needs most of the 180 GB at N = 2
        del a, a_nz, K, f, part
        torch.cuda.empty_cache()
        c4n = int(os.environ.get("FB2_C4_NEL", "320"))      # debugging knob: a smaller total size
        config4 = config4_strong(fb, ctx, dist, torch, world, rank, dev, steps=max(3, min(args.steps, 5)), nel_total=(c4n,) * 3)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = load_peaks()
    fp64_peak = ctx.measure_fp64_peak()
    cells_per_s_kernel = gcells0 / (kernel_ms * 1e-3)
    achieved = cfg["bmin"] * cells_per_s_kernel / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.config)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms, "bytes_per_cell": cfg["bmin"],
                "kernel_cells": gcells0,
                "fp64": {"achieved_tflops": cfg["fmin"] * cells_per_s_kernel / 1e12, "peak_tflops": fp64_peak,
                         "frac": cfg["fmin"] * cells_per_s_kernel / 1e12 / fp64_peak if fp64_peak else None,
                         "flop_per_cell": cfg["fmin"], "peak_source": "FMA microbenchmark in this run"}}
    if world == 1:
        par = "1 GPU"
    else:
        par = (f"{world} ranks, {'x'.join(map(str, dims))} blocks of {'x'.join(map(str, nel))} cells; "
               + ("own cells + NCCL exchange of interface columns" if args.dist_mode == "exchange"
                  else "own + halo cells, no communication"))
    line = {
        "metric": "cells assembled/sec (K+f, FP64, 3D hex)", "value": value, "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"] + ("" if world == 1 else f" per GPU, {ncells_total} cells in total"),
                   "cells": ncells_total, "ndofs_rank0": ndofs0, "nnz_rank0": nnz0, "scatter": args.scatter,
                   "l2": "inputs+outputs (>= 3 GB per GPU) exceed the 126 MB L2; no flush needed", "setup_s": round(t_setup, 2),
                   "parallelism": par},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "checks": checks,
    }
    line["checks"].update(nccl_check)
    if config4:
        line["config4"] = config4
    if e2e:
        line["e2e"] = e2e
    if apply_info:
        line["apply"] = apply_info
    if spmv:
        spmv["frac_of_hbm_peak"] = spmv["GB/s"] / hbm_peak
        line["spmv"] = spmv
    if ea_info:
        if "mul_GB/s" in ea_info:
            ea_info["mul_frac_of_hbm_peak"] = ea_info["mul_GB/s"] / hbm_peak
        line["element_assembly"] = ea_info
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, SAMPLE(cfg))
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

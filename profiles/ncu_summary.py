#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + top stall instructions (needs `ncu` on PATH, no GPU)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def main(rep, ntop=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?")[:120])
        for k in WANT:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    data = []
    for n, r in enumerate(rows[hi + 1:]):
        try:
            data.append((float(r[ci["# Samples"]]), n, r))
        except (ValueError, IndexError):
            pass
    tot = sum(v for v, _, _ in data) or 1
    print(f"top stall instructions ({len(data)} SASS instructions, {tot:.0f} samples):")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for v, n, r in sorted(data, key=lambda t: -t[0])[:ntop]:
        top = max(stall_cols, key=lambda h: float(r[ci[h]] or 0))
        print(f"  #{n:5d} {100 * v / tot:5.1f}%  {top:18s} {r[ci['Source']].strip()[:80]}")
    agg = {h: sum(float(r[ci[h]] or 0) for _, _, r in data) for h in stall_cols}
    print("stall totals:", ", ".join(f"{h[6:]}={100 * x / tot:.1f}%" for h, x in sorted(agg.items(), key=lambda t: -t[1])[:8]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

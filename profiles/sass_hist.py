#!/usr/bin/env python
"""Executed-instruction histogram of an .ncu-rep by SASS address ranges: splits the kernel at the given instruction
indices (or every N instructions) and prints executed warp instructions, stall samples and the opcode mix per range.
usage: sass_hist.py report.ncu-rep [idx0 idx1 ...]"""
import csv, io, subprocess, sys, collections

def main(rep, cuts):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) > ci["Instructions Executed"]]
    tot_i = sum(float(r[ci["Instructions Executed"]] or 0) for r in data)
    tot_s = sum(float(r[ci["# Samples"]] or 0) for r in data)
    if not cuts:
        cuts = list(range(0, len(data), 100))
    cuts = sorted(set([0] + cuts + [len(data)]))
    for a, b in zip(cuts[:-1], cuts[1:]):
        seg = data[a:b]
        ex = sum(float(r[ci["Instructions Executed"]] or 0) for r in seg)
        sm = sum(float(r[ci["# Samples"]] or 0) for r in seg)
        ops = collections.Counter()
        for r in seg:
            t = r[ci["Source"]].split()
            op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
            ops[op.split(".")[0]] += float(r[ci["Instructions Executed"]] or 0)
        top = ", ".join(f"{k}:{100*v/ex:.0f}%" for k, v in ops.most_common(6)) if ex else ""
        print(f"[{a:5d},{b:5d}) exec {100*ex/tot_i:5.1f}%  samples {100*sm/tot_s:5.1f}%  per-launch-warp-instr {ex:.3e}  {top}")

if __name__ == "__main__":
    main(sys.argv[1], [int(x) for x in sys.argv[2:]])

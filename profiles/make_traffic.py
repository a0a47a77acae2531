#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list with DRAM metrics (the `--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` pass of scripts/r2_final_profile.sh): DRAM bytes per launch of the dominant kernel of each config.
usage: make_traffic.py config=launches.csv:kernel_regex ..."""
import csv, json, os, re, sys
from collections import OrderedDict

out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    cfg, rest = arg.split("=", 1)
    path, rx = rest.rsplit(":", 1)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = OrderedDict()
    for r in rows[1:]:
        if re.search(rx, r[ki]):
            per.setdefault(r[ii], {})[r[mi]] = float(r[vi].replace(",", ""))
    vals = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values() if "dram__bytes_read.sum" in d]
    vals = vals[:max(1, (len(vals) + 1) // 2)]       # warm-up + timed steps come first (zero fill + kernel); the later launches are
                                                      # bench.py's kernel-only timing in accumulate mode (every column reduce-added)
    out[cfg] = int(sorted(vals)[len(vals) // 2])
    print(cfg, out[cfg], f"({len(vals)} launches of /{rx}/)")
json.dump(out, open(out_path, "w"), indent=1)

"""Bucket an ncu report's SASS by the regions between CTA barriers: share of executed instructions and stall samples.
usage: python profiles/sass_regions.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
tot = sum(int(r[ix]) for r in data); ts = sum(int(r[isamp]) for r in data)
print(rows[0][1], "instructions", tot, "samples", ts)
acc = [[0, 0, 0, 0, ""]]
for i, r in enumerate(data):
    s = r[1].strip()
    acc[-1][0] += int(r[ix]); acc[-1][1] += int(r[isamp]); acc[-1][2] += 1
    if "DMMA" in s: acc[-1][3] += 1
    if "BAR.SYNC" in s or "WARPSYNC.ALL" in s and len(sys.argv) > 2:
        acc[-1][4] = f"ends at #{i} {s}"
        acc.append([0, 0, 0, 0, ""])
for a in acc:
    print(f"{a[0]/tot*100:5.1f}% inst {a[1]/ts*100:5.1f}% samples  nsass={a[2]} dmma={a[3]} {a[4]}")

# round-2 final record of the Q1^3 elasticity path: bench lines (new default kernel and the warp-per-cell kernel), launch list, full ncu capture
mkdir -p gpurun_out
for cfg in c5 c5full; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_${cfg}_n1.json 2> gpurun_out/r02_bench_${cfg}.err; tail -2 gpurun_out/r02_bench_${cfg}.err
done
timeout 300 python bench.py --config c5 --variant 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_syrk_n1.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
bash scripts/r2_vec_prof.sh r02_prof_c5_final
cuobjdump -sass ferrite.jl_b200/lib/libferrite_b200.so 2>/dev/null | awk '/Function : .*k_march_vecILb0ELb1/{p=1} p&&/Function : /&&!/k_march_vecILb0ELb1/{p=0} p' | grep -E "DMMA|UBLKCP|UBLKRED|REDG|LDGSTS|BAR.SYNC" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//' | sed 's/ *\/\*.*//' | sort | uniq -c | sort -rn | head -20 > gpurun_out/r02_sass_k_march_vec.txt
python - <<'PY' > gpurun_out/r02_sass_k_march_hex.txt
import subprocess, re, collections
out = subprocess.run("cuobjdump -sass ferrite.jl_b200/lib/libferrite_b200.so", shell=True, capture_output=True, text=True).stdout
on = False; ops = []
for line in out.splitlines():
    if "Function :" in line:
        on = "k_march_hexILi1ELb0ELb1" in line
        continue
    if on:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m: ops.append(m.group(1))
h = collections.Counter(o.split(".")[0] for o in ops)
print("SASS of k_march_hex<FB2_ELEM_HEAT, CHECK=false, ANALYTIC=true> (sm_100a, cuobjdump -sass of libferrite_b200.so, round 2, final)")
print(f"{len(ops)} instructions; opcode histogram (static):")
for k, v in h.most_common(40): print(f"  {k:10s} {v}")
print("bulk (TMA) and tensor forms:")
for k, v in collections.Counter(o for o in ops if o.startswith(("UBLK", "DMMA", "REDG", "LDGSTS", "UTMA"))).items(): print(f"  {k:40s} {v}")
PY

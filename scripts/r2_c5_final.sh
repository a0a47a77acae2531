# round-2 final record of the Q1^3 elasticity path: bench lines (new default kernel and the warp-per-cell kernel), launch list, full ncu capture
mkdir -p gpurun_out
for cfg in c5 c5full; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_${cfg}_n1.json 2> gpurun_out/r02_bench_${cfg}.err; tail -2 gpurun_out/r02_bench_${cfg}.err
done
timeout 300 python bench.py --config c5 --variant 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_syrk_n1.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
bash scripts/r2_vec_prof.sh r02_prof_c5_final
cuobjdump -sass ferrite.jl_b200/lib/libferrite_b200.so 2>/dev/null | awk '/Function : .*k_march_vecILb0ELb1/{p=1} p&&/Function : /&&!/k_march_vecILb0ELb1/{p=0} p' | grep -E "DMMA|UBLKCP|UBLKRED|REDG|LDGSTS|BAR.SYNC" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//' | sed 's/ *\/\*.*//' | sort | uniq -c | sort -rn | head -20 > gpurun_out/r02_sass_k_march_vec.txt

# full ncu capture of k_march_vec on C5 (128^3 Q1^3 elasticity) + summaries
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_march_vec -s 4 -c 1 -f -o /tmp/prof_c5 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1
tail -1 gpurun_out/ncu_run.log
(python profiles/ncu_summary.py /tmp/prof_c5.ncu-rep 40; python profiles/sass_hist.py /tmp/prof_c5.ncu-rep) > gpurun_out/${1:-r02_prof_c5_march_a}.txt 2>&1
cp /tmp/prof_c5.ncu-rep gpurun_out/prof_c5.ncu-rep

mkdir -p gpurun_out
python scripts/c2_variants.py 0 31 30 2>&1 | tail -1 | tee gpurun_out/r2_c2_variants_d.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or large_entrywise or assembly_matches" 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_march_hex -s 3 -c 1 -f -o /tmp/prof_c2 python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
(python profiles/ncu_summary.py /tmp/prof_c2.ncu-rep 30; python profiles/sass_hist.py /tmp/prof_c2.ncu-rep) > gpurun_out/r02_prof_c2_march_d.txt 2>&1
cp /tmp/prof_c2.ncu-rep gpurun_out/r02_c2_march_d.ncu-rep

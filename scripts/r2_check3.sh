mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or assembly_matches_oracle or fillzero or host_buffer" 2>&1 | tail -5
python scripts/c2_variants.py 0 31 30 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c2_c.json 2> gpurun_out/r2_bench_c2_c.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_c.json')); print('c2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_march_hex -s 3 -c 1 -f -o /tmp/prof_c2 python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
(python profiles/ncu_summary.py /tmp/prof_c2.ncu-rep 30; python profiles/sass_hist.py /tmp/prof_c2.ncu-rep) > gpurun_out/r02_prof_c2_march_b.txt 2>&1
cp /tmp/prof_c2.ncu-rep gpurun_out/r02_c2_march_b.ncu-rep

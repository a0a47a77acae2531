mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0; do timeout 120 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e --variant $v > gpurun_out/bench_c2_v$v.json 2> gpurun_out/bench_c2_v$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_c2_v$v.json')); print('c2', $v, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done

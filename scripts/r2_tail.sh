mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or incomplete or entrywise or large or partitioned" 2>&1 | tail -3
VS_KIND=heat timeout 300 python scripts/vec_sizes.py 200x200x200 201x200x200 201x201x201 64x64x64 2>&1 | grep -o '"nel.*"ms": [0-9.]*'
timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_tail_c2.json 2> gpurun_out/r2_tail_c2.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_tail_c2.json')); print('c2', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"
for r in 0 1; do PL_KIND=heat PL_NEL=200 timeout 600 python scripts/part_local_time.py 2 $r 2>&1 | tail -1 | cut -c1-330; done

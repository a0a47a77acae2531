mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or partitioned_assembly or entrywise or large" 2>&1 | tail -3
for cfg in c2 c5; do
timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_fl_$cfg.json 2> gpurun_out/r2_fl_$cfg.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_fl_$cfg.json')); print('$cfg', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"; tail -2 gpurun_out/r2_fl_$cfg.err
done

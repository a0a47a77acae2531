mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or partitioned_assembly or incomplete or hex-q1-elast or streamed" 2>&1 | tail -3
for z in 1 0; do
FB2_MARCH_ZSEL=$z timeout 300 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_zs_$z.json 2> gpurun_out/r2_zs_$z.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_zs_$z.json')); print('zsel$z', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"; tail -2 gpurun_out/r2_zs_$z.err
done

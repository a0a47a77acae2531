mkdir -p gpurun_out
FB2_MARCH_ZSEL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or incomplete or entrywise or large" 2>&1 | tail -3
for z in 1 0; do
FB2_MARCH_ZSEL=$z timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_zs2_$z.json 2> gpurun_out/r2_zs2_$z.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_zs2_$z.json')); print('c2 zsel$z', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"; tail -2 gpurun_out/r2_zs2_$z.err
done
timeout 300 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_zs2_c5.json 2> gpurun_out/r2_zs2_c5.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_zs2_c5.json')); print('c5', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"

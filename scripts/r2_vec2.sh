mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching_tile_kernel_elasticity" 2>&1 | tail -5
FB2_MVEC_FLUSH=thread timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching_tile_kernel_elasticity" 2>&1 | tail -3
for mode in tma thread; do
FB2_MVEC_FLUSH=$mode timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_vec_c5_$mode.json 2> gpurun_out/r2_vec_c5_$mode.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_vec_c5_$mode.json')); print('$mode', d['ms_per_step'], d['roofline']['kernel_ms'], d['checks'])"; tail -3 gpurun_out/r2_vec_c5_$mode.err
done

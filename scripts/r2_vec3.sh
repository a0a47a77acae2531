mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching_tile_kernel_elasticity" 2>&1 | tail -3
for dbg in ${DBGS:-0 1 3}; do
FB2_MVEC_DBG=$dbg timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_vec_c5_dbg$dbg.json 2> gpurun_out/r2_vec_c5_dbg$dbg.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_vec_c5_dbg$dbg.json')); print('dbg$dbg', d['ms_per_step'], d['roofline']['kernel_ms'])"; tail -3 gpurun_out/r2_vec_c5_dbg$dbg.err
done

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching_tile_kernel_elasticity" 2>&1 | tail -15
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_vec_c5.json 2> gpurun_out/r2_vec_c5.err; cut -c1-1500 gpurun_out/r2_vec_c5.json; tail -3 gpurun_out/r2_vec_c5.err
timeout 300 python bench.py --config c5 --variant 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_vec_c5_v32.json 2> gpurun_out/r2_vec_c5_v32.err; cut -c1-600 gpurun_out/r2_vec_c5_v32.json; tail -3 gpurun_out/r2_vec_c5_v32.err

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching_tile_kernel_elasticity or partitioned_assembly" 2>&1 | tail -3
timeout 800 python scripts/vec_sizes.py 128x128x128 129x128x128 127x128x128 128x129x128 160x160x160 161x160x160 2>&1 | tail -6

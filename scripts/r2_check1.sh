# round 2, first GPU check of the marching-tile kernel: parity tests, then C2 timing old vs new and a chunk-length sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marching or assembly_matches_oracle or fillzero or host_buffer" 2>&1 | tail -8
python scripts/c2_variants.py 0 30 > gpurun_out/r2_c2_variants.txt 2>&1; cat gpurun_out/r2_c2_variants.txt
for lz in 8 12 25 50 200; do echo "lz=$lz"; FB2_MARCH_LZ=$lz python scripts/c2_variants.py 0 2>&1 | tail -1; done | tee gpurun_out/r2_c2_lz.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err; cat gpurun_out/r2_bench_c2_a.json

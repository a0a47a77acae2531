mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python scripts/c2_variants.py 0 30 2>&1 | tail -1 | tee gpurun_out/r2_c2_variants_f.txt
FB2_MARCH_ZSEL=0 python scripts/c2_variants.py 0 2>&1 | tail -1
for lz in 13 25 34; do echo "lz=$lz"; FB2_MARCH_LZ=$lz python scripts/c2_variants.py 0 2>&1 | tail -1; done | tee gpurun_out/r2_c2_lz3.txt

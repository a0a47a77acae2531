mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python scripts/c2_variants.py 0 2>&1 | tail -1

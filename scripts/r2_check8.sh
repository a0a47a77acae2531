python scripts/c2_variants.py 0 2>&1 | tail -1
python scripts/part_local_time.py 2 1 2>&1 | tail -1
python scripts/part_local_time.py 8 0 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4

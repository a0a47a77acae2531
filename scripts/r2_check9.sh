python scripts/c2_variants.py 0 2>&1 | tail -1
for zc in 0 16 64; do echo "zc=$zc"; FB2_MARCH_ZC=$zc python scripts/c2_variants.py 0 2>&1 | tail -1; done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3

for r in 0 2 4 6 8; do
echo "rpad $r: $(FB2_MARCH_RPAD=$r VS_KIND=heat timeout 300 python scripts/vec_sizes.py 200x200x200 201x200x200 2>&1 | grep -o '"ms": [0-9.]*' | tr '\n' ' ')"
done

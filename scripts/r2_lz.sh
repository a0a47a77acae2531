for lz in 0 8 12 16 32 64; do
if [ $lz = 0 ]; then unset FB2_MARCH_LZ; else export FB2_MARCH_LZ=$lz; fi
echo "${VS_KIND:-elasticity} lz $lz: $(timeout 300 python scripts/vec_sizes.py ${SZ:-128x128x128} 2>&1 | grep -o '"ms": [0-9.]*')"
done

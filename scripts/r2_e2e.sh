for ns in 8 16 20 4; do
FB2_HOST_SLABS=$ns timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_e2e_$ns.json 2> gpurun_out/r2_e2e_$ns.err; python -c "
import json; d=json.load(open('gpurun_out/r2_e2e_$ns.json')); print('slabs $ns', d['e2e']['value'], 8e6/d['e2e']['value']*1e3, 'ms')"
done

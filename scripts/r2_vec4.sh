mkdir -p gpurun_out
for cfg in c5 c5full; do for v in 0 32; do
timeout 300 python bench.py --config $cfg --variant $v --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_vec_${cfg}_v$v.json 2> gpurun_out/r2_vec_${cfg}_v$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_vec_${cfg}_v$v.json')); print('$cfg v$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['setup_s'])"; tail -3 gpurun_out/r2_vec_${cfg}_v$v.err
done; done

mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in c2 c5 c3 c4; do for v in 0 1; do if [ $v = 1 ] && [ $c = c2 -o $c = c4 ]; then continue; fi
timeout 200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e --variant $v > gpurun_out/bench_${c}_v$v.json 2> gpurun_out/bench_${c}_v$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${c}_v$v.json')); print('$c', $v, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done; done

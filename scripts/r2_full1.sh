mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_ref_n1.json 2> gpurun_out/r2_ref_n1.err; cut -c1-400 gpurun_out/r2_ref_n1.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

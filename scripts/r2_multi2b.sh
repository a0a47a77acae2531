# 2-GPU check: emulated + real NCCL partition tests, then the bench at N=2 with configs[4] at 160^3 per rank
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu -k "nccl or partition" 2>&1 | tail -5
FB2_C4_NEL=${C4NEL:-200} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 3000 gpurun_out/r2_bench_n2.json; grep -i -E "error|Traceback|NCCL INFO comm|nranks" gpurun_out/r2_bench_n2.err | head -10

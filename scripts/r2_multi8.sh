mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc=$?"
tail -c 4500 gpurun_out/r2_bench_n$N.json; grep -i -E "Traceback|Error|nranks|NCCL INFO comm 0x" gpurun_out/r2_bench_n$N.err | head -8

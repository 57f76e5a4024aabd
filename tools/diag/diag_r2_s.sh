#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s_bench_n8.json 2> gpurun_out/s_bench_n8.err; echo "bench n8 rc $?"
tail -c 3000 gpurun_out/s_bench_n8.json
grep -i "NCCL INFO.*\(NVLS\|Connected\|comm 0x\|nranks\)" gpurun_out/s_bench_n8.err | head -12
tail -5 gpurun_out/s_bench_n8.err

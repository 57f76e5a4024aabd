#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
env | grep -i nccl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 --steps 3 --warmup 3 --train 0 --alt-precisions 0 > gpurun_out/nccl_check.json 2> gpurun_out/nccl_check.err; echo "rc $?"
grep -c "NCCL INFO" gpurun_out/nccl_check.err; grep "NCCL INFO" gpurun_out/nccl_check.err | grep -i "nranks\|Init COMPLETE\|NVLS\|P2P" | head -6
python -c "import json; j=json.loads(open('gpurun_out/nccl_check.json').read().strip().splitlines()[-1]); print(j['value'], j['n_gpus'])"

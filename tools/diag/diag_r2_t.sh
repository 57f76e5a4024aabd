#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for v in "" "--eager-backbone"; do
  timeout 300 python tools/train_step.py --steps 10 --warmup 3 $v 2> gpurun_out/t_n1.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N1 [$v]', round(j['ms_per_step'],2), j['segments_ms'])" || tail -5 gpurun_out/t_n1.err
done
for v in "" "--eager-backbone"; do
  timeout 300 $TR tools/train_step.py --sync-bn --steps 10 --warmup 3 $v 2> gpurun_out/t_n2.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N2 [$v]', round(j['ms_per_step'],2), j['segments_ms'], j['peer_exchanges_per_step'], j['loss'])" || tail -5 gpurun_out/t_n2.err
done
timeout 300 $TR tools/train_step.py --sync-bn --local-backbone-bn --steps 10 --warmup 3 2> gpurun_out/t_n2b.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N2 local-bb-bn', round(j['ms_per_step'],2), j['segments_ms'], j['peer_exchanges_per_step'])" || tail -5 gpurun_out/t_n2b.err
timeout 600 python -m pytest tests -m gpu -x -q -k "peer or ddp or two_gpu or 2gpu or reducer or head or psm" > gpurun_out/t_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/t_pytest.log
timeout 600 python bench.py --no-cpu-baseline --train 0 --ops 0 --gpu-torch-baseline 0 --alt-precisions 0 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; python -c "
import json; j=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['segments_ms'])"

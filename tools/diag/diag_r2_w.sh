#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "graphs or peer or data_parallel" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/w_pytest.log
for v in "" "--graph-hot-path"; do
  timeout 300 $TR tools/train_step.py --sync-bn --steps 10 --warmup 3 $v 2> gpurun_out/w_n2.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N2 [$v]', round(j['ms_per_step'],2), j['segments_ms'], j['peer_exchanges_per_step'], j['loss'])" || tail -25 gpurun_out/w_n2.err
done

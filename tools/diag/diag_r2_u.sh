#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in "" "--graph-hot-path"; do
  timeout 400 python tools/train_step.py --steps 10 --warmup 3 $v 2> gpurun_out/u_n1.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N1 [$v]', round(j['ms_per_step'],2), j['segments_ms'], j['loss'], j['peak_mem_gb'])" || tail -25 gpurun_out/u_n1.err
done
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/u_pytest.log

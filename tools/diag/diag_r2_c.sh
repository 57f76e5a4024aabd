#!/usr/bin/env bash
# round-2 call C (1 GPU): tcgen05 weight gradient -- tests first (short timeout: a new kernel), then the suite, then timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -s -k "wgrad_tcgen05" --timeout 120 > gpurun_out/pytest_wgrad_c.log 2>&1; echo "wgrad tests=$?"; tail -25 gpurun_out/pytest_wgrad_c.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest=$?"; tail -12 gpurun_out/pytest_gpu_c.log
echo "== train, tc wgrad"
timeout 300 python tools/train_step.py --steps 4 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_c_tc.json
echo "== train, simt wgrad"
DMB_B200_TRAIN_TC_WGRAD=0 timeout 300 python tools/train_step.py --steps 4 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_c_simt.json
echo "== launches of one training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file gpurun_out/launches_train_c.csv python tools/train_step.py --steps 1 --warmup 2 > gpurun_out/train_ncu_c.log 2>&1; echo "ncu=$?"

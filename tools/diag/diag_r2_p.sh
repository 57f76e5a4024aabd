#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
P=tools/_build/tma_f32_probe
{
for args in "4 36 -2 40" "4 36 0 40" "4 36 -4 40" "4 32 -2 40" "4 32 0 40" "4 64 -2 40" "4 40 -2 40" "3 36 -2 40" "3 32 0 40" "4 36 -2 64" "4 36 30 64" "4 36 62 64"; do
  timeout 60 $P $args
done
} > gpurun_out/p_probe.log 2>&1
cat gpurun_out/p_probe.log
export DMB_B200_LGA_ROT=2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/p_pytest.log
timeout 600 python tools/train_step.py --steps 10 --warmup 3 > gpurun_out/p_train.json 2> gpurun_out/p_train.err; tail -3 gpurun_out/p_train.json

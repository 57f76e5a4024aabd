#!/usr/bin/env bash
# round-2 call D (1 GPU): transposed conv class-group kernels (kind 6): tests, A/B bench, launch lists; training with channels_last backbone
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "strided or hourglass or config1" --timeout 120 > gpurun_out/pytest_tr_d.log 2>&1; echo "transposed tests=$?"; tail -6 gpurun_out/pytest_tr_d.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest=$?"; tail -6 gpurun_out/pytest_gpu_d.log
FAST="--train 0 --ops 0 --gpu-torch-baseline 0 --no-cpu-baseline --alt-precisions 0 --steps 20"
for g in 1 0 1 0; do
  DMB_B200_TC_DECONV_GROUPS=$g timeout 300 python bench.py $FAST > gpurun_out/bench_d_groups$g.json 2>> gpurun_out/bench_d.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_d_groups$g.json"))
print("groups=$g pairs/s %.1f ms/step %.2f segments %s frac %.3f" % (d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"]))
PY
done
echo "== hot path launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc|head_gather|upsample|cat_volume" -c 300 --csv \
    --log-file gpurun_out/launches_hot_path_d.csv python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_d.log 2>&1; echo "ncu launches=$?"
echo "== train channels_last / nchw backbone"
timeout 300 python tools/train_step.py --steps 4 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_d_cl.json
timeout 300 python tools/train_step.py --steps 4 --warmup 2 --nchw-backbone 2>&1 | tail -1 | tee gpurun_out/train_d_nchw.json
echo "== our kernels in one training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dmb|conv3d_tc" -s 1200 -c 700 --csv --log-file gpurun_out/launches_train_d.csv python tools/train_step.py --steps 1 --warmup 2 > gpurun_out/train_ncu_d.log 2>&1; echo "ncu=$?"

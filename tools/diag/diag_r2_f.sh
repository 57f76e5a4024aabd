#!/usr/bin/env bash
# round-2 call F (1 GPU): stride-2 tcgen05 wgrad + padded head wgrad, rewritten cat/dif/gwc volume kernels
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -q -x -k "wgrad or volume or gwc or cat_fms or dif" --timeout 200 > gpurun_out/pytest_f1.log 2>&1; echo "new-kernel tests=$?"; tail -8 gpurun_out/pytest_f1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest=$?"; tail -6 gpurun_out/pytest_gpu_f.log
timeout 300 python tools/train_step.py --steps 4 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_f.json
timeout 600 python bench.py --no-cpu-baseline --alt-precisions 0 --gpu-torch-baseline 0 --train 0 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_f.json"))
print("pairs/s %.1f ms/step %.2f segments %s frac %.3f cat frac %.3f" % (d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"], d["roofline_cat_volume"]["frac"]))
print(json.dumps(d["ops"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bn_|conv3d|wgrad|blocked|ncs_|up8|upsample|focal|soft_argmin|cat_volume|head_gather|pack_weights" -s 1120 -c 600 --csv --log-file gpurun_out/launches_train_f.csv python tools/train_step.py --steps 1 --warmup 2 > gpurun_out/train_ncu_f.log 2>&1; echo "ncu=$?"

#!/usr/bin/env bash
# round-2 call E (1 GPU): the library's kernels inside one training step (launch list), fill bandwidth
set -u
mkdir -p gpurun_out
timeout 60 python tools/peak_write.py | tee gpurun_out/peak_write.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bn_|conv3d|wgrad|blocked|ncs_|up8|upsample|focal|soft_argmin|cat_volume|head_gather|pack_weights" -s 1060 -c 560 --csv --log-file gpurun_out/launches_train_e.csv python tools/train_step.py --steps 1 --warmup 2 > gpurun_out/train_ncu_e.log 2>&1; echo "ncu=$?"
tail -3 gpurun_out/launches_train_e.csv | cut -c1-300

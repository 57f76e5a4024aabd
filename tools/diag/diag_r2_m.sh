#!/usr/bin/env bash
# round-2 call M (1 GPU): ncu --set full captures: dominant trunk kernel, LGA, GWC, cat volume
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"lga_r2_rot|gwc_rows_unit|cat_volume_blocked|volume_rows" -s 4 -c 4 -o gpurun_out/r2_prof_ops python tools/profile_ops.py > gpurun_out/prof_ops_m.log 2>&1; echo "ncu ops=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 79 -c 1 -o gpurun_out/r2_prof_k3 python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_k3_m.log 2>&1; echo "ncu k3=$?"
ls -la gpurun_out/*.ncu-rep

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 103 -c 1 -f -o gpurun_out/r2_prof_k7 python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_k7.log 2>&1; echo "ncu k7=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 151 -c 1 -f -o gpurun_out/r2_prof_head python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_head.log 2>&1; echo "ncu head=$?"
ls -la gpurun_out/*.ncu-rep

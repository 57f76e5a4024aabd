#!/usr/bin/env bash
# round-2 call N (1 GPU): TMA-staged LGA, GWC epilogue; full suite; ops timings
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lga or ganet or config4 or gwc or config3 or correlation1d" --timeout 300 > gpurun_out/pytest_n1.log 2>&1; echo "lga/gwc tests=$?"; tail -4 gpurun_out/pytest_n1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest=$?"; tail -4 gpurun_out/pytest_gpu_n.log
for v in 1 2; do
  DMB_B200_LGA_ROT=$v timeout 300 python - <<PY
import json, sys, torch
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench, bench_blocks
r = bench_blocks.ops_block(torch.device("cuda", 0), bench.peaks()["hbm_gbs"])
print("LGA_ROT=$v", json.dumps({k.split(" ")[0]: (v["ms"], v["frac_of_hbm_peak"]) for k, v in r.items()}))
PY
done

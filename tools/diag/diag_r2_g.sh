#!/usr/bin/env bash
# round-2 call G (1 GPU): bidirectional / L2-blocked SGA, rotating-accumulator LGA
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sga or lga or ganet or config4" --timeout 300 > gpurun_out/pytest_g1.log 2>&1; echo "sga/lga tests=$?"; tail -8 gpurun_out/pytest_g1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu_g.log
for v in "1 1" "0 1" "1 0"; do
  set -- $v
  DMB_B200_SGA_BIDIR=$1 DMB_B200_LGA_ROT=$2 timeout 300 python - <<PY
import json, sys, torch
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench, bench_blocks
r = bench_blocks.ops_block(torch.device("cuda", 0), bench.peaks()["hbm_gbs"])
print("SGA_BIDIR=$1 LGA_ROT=$2", json.dumps({k.split(" ")[0]: (v["ms"], v["frac_of_hbm_peak"]) for k, v in r.items()}))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"sga|lga" -c 60 --csv --log-file gpurun_out/ncu_ganet_g.csv python tools/profile_ganet.py > gpurun_out/prof_ganet_g.log 2>&1; echo "ncu=$?"

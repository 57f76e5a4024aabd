#!/usr/bin/env bash
# round-2 call H (1 GPU): head Q-planes (9 instead of 27 tap planes), SGA prefetch depth, full suite, bench A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or medium or acfnet or sga or golden" --timeout 300 > gpurun_out/pytest_h1.log 2>&1; echo "head/sga tests=$?"; tail -6 gpurun_out/pytest_h1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu_h.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
for v in "1" "0"; do
  DMB_B200_SGA_BIDIR=$v timeout 300 python - <<PY
import json, sys, torch
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench, bench_blocks
r = bench_blocks.ops_block(torch.device("cuda", 0), bench.peaks()["hbm_gbs"])
print("SGA_BIDIR=$v", json.dumps({k.split(" ")[0]: (v["ms"], v["frac_of_hbm_peak"]) for k, v in r.items()}))
PY
done
FAST="--train 0 --ops 0 --gpu-torch-baseline 0 --no-cpu-baseline --alt-precisions 0 --steps 20"
timeout 300 python bench.py $FAST > gpurun_out/bench_h.json 2>> gpurun_out/bench_h.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_h.json"))
print("pairs/s %.1f ms/step %.2f segments %s frac %.3f" % (d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc|head_gather|upsample|cat_volume" -c 300 --csv \
    --log-file gpurun_out/launches_hot_path_h.csv python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_h.log 2>&1; echo "ncu launches=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"sga" -c 60 --csv --log-file gpurun_out/ncu_sga_h.csv python tools/profile_ganet.py > gpurun_out/prof_sga_h.log 2>&1; echo "ncu=$?"

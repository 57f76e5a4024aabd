#!/usr/bin/env bash
# One measurement bundle per gpurun call (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_bundle.sh all'
# Sections (any subset as arguments): tests  smoke  bench  ops  train  launches  full  probes
# Everything lands in gpurun_out/ (merged back by gpurun); numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
SECTIONS=("$@")
[[ ${#SECTIONS[@]} -eq 0 ]] && SECTIONS=(tests smoke bench)
has() { for s in "${SECTIONS[@]}"; do [[ "$s" == "$1" || "$s" == "all" ]] && return 0; done; return 1; }

if has tests; then
    timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/pytest_gpu.log
fi
if has smoke; then
    timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
fi
if has bench; then
    timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench=$?"
    python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print("pairs/s %.1f  ms/step %.2f  e2e %.1f  segments %s  roofline.frac %.3f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"]))
PY
fi
if has ops; then
    timeout 400 python tools/bench_ops.py > gpurun_out/bench_ops.log 2>&1; tail -1 gpurun_out/bench_ops.log | head -c 400; echo
fi
if has train; then
    timeout 300 python tools/train_step.py --steps 5 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_n1.json
fi
if has launches; then      # per-launch device times of the warm hot-path pass
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc|head_gather|upsample|cat_volume" -c 300 --csv \
        --log-file gpurun_out/launches_hot_path.csv python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof.log 2>&1; echo "ncu launches=$?"
fi
if has full; then          # one full capture of the dominant kernel (32->32 stride-1 layer of the second pass)
    ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 70 -c 1 -o gpurun_out/prof_k3n4 \
        python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_full.log 2>&1; echo "ncu full=$?"
fi
if has probes; then        # micro-benchmarks behind DESIGN.md section 5 (binaries built by the nvcc lines in their headers)
    [[ -x tools/_build/mma_probe ]] && timeout 60 tools/_build/mma_probe | head -8
    [[ -x tools/_build/mma_mn_probe ]] && timeout 60 tools/_build/mma_mn_probe
    [[ -x tools/_build/wgrad_tc_proto ]] && { timeout 60 tools/_build/wgrad_tc_proto 0 1; timeout 60 tools/_build/wgrad_tc_proto 1 1; }
    timeout 120 python tools/tc_trace.py | head -8
    timeout 200 python tools/tc_clock.py
fi

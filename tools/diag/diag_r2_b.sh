#!/usr/bin/env bash
# round-2 call B (1 GPU): full GPU test suite, smoke, bench with the new blocks
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest=$?"; tail -15 gpurun_out/pytest_gpu_b.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench=$?"; tail -5 gpurun_out/bench_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_b.json"))
print("pairs/s %.1f  ms/step %.2f  e2e %.1f  segments %s  roofline.frac %.3f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"]))
print("train", json.dumps(d.get("train"))[:700])
print("ops", json.dumps(d.get("ops")))
print("gpu_torch", json.dumps(d.get("gpu_torch_baseline")))
print("cpu", json.dumps(d.get("cpu_baseline")))
PY

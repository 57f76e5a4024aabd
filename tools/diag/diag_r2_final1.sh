#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/final_pytest.log
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/final_smoke.log
timeout 1200 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "bench ref rc $?"; tail -c 600 gpurun_out/final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline --alt-precisions 0 --train 0 --ops 0 --gpu-torch-baseline 0 > gpurun_out/final_ncu_bench.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/final_bench_n1.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step", "segments_ms", "gpu_launches")})
print("e2e", j["e2e"]["value"], "roofline", j["roofline"]["frac"], "cpu", j.get("cpu_baseline", {}).get("value"))
print("train", j["train"]["ms_per_step"], j["train"]["segments_ms"])
print("ops", json.dumps(j["ops"])[:1500])
print("torch", json.dumps(j["gpu_torch_baseline"])[:600])
PY

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r_smoke.log 2>&1; rc=$?; echo "smoke rc $rc"; tail -3 gpurun_out/r_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q -k "psm or acf or head or aggreg or full_size or config2" > gpurun_out/r_pytest_head.log 2>&1; echo "pytest(head subset) rc $?"; tail -4 gpurun_out/r_pytest_head.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc $?"
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r_bench.json").read().strip().splitlines()[-1])
    print({k: j[k] for k in ("value", "ms_per_step", "e2e", "roofline") if k in j})
    print(j.get("segments_ms") or j.get("config"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r_bench.err").read()[-2000:])
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r_pytest.log

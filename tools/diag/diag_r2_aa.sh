#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aa_smoke.log 2>&1; rc=$?; echo "smoke rc $rc"; tail -2 gpurun_out/aa_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
for m in 1 0 1; do
DMB_B200_TC_BALANCED=$m timeout 600 python bench.py --no-cpu-baseline --train 0 --ops 0 --gpu-torch-baseline 0 --alt-precisions 0 > gpurun_out/aa_bench_$m.json 2> gpurun_out/aa_bench_$m.err; python -c "
import json; j=json.loads(open('gpurun_out/aa_bench_$m.json').read().strip().splitlines()[-1]); print('BALANCED=$m', j['value'], j['segments_ms'], j['clocks']['sm_mhz'])"
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/aa_pytest.log

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "upsample or regress or argmin or psm or full_size or lga" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/y_pytest.log
for m in 0 1; do
DMB_B200_REGRESS_IV=$m timeout 600 python bench.py --no-cpu-baseline --train 0 --ops 0 --gpu-torch-baseline 0 --alt-precisions 0 > gpurun_out/y_bench_$m.json 2> gpurun_out/y_bench_$m.err; python -c "
import json; j=json.loads(open('gpurun_out/y_bench_$m.json').read().strip().splitlines()[-1]); print('IV=$m', j['value'], j['segments_ms'])"
done

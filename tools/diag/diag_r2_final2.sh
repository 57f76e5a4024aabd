#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench n$N rc $?"
python - <<PY
import json
j = json.loads(open("gpurun_out/final_bench_n$N.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "n_gpus", "ms_per_step")}, "e2e", j["e2e"]["value"])
t = j["train"]; print("train", t["ms_per_step"], t["segments_ms"], t["peer_exchanges_per_step"]); print("variant", t.get("variant_backbone_bn_per_rank"))
PY
grep -c "NCCL INFO" gpurun_out/final_bench_n$N.err; grep -m3 "NCCL INFO.*\(Init COMPLETE\|nranks\)" gpurun_out/final_bench_n$N.err

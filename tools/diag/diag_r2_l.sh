#!/usr/bin/env bash
# round-2 call L (2 GPUs): 2-GPU tests, bench under torchrun with default env and with NCCL_DEBUG=INFO (stdout must be one JSON line)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_autograd_paths.py -m gpu -q -k "two_gpus or tensor_device" --timeout 600 > gpurun_out/pytest_l1.log 2>&1; echo "2-gpu tests=$?"; tail -4 gpurun_out/pytest_l1.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_l_n2.json 2> gpurun_out/bench_l_n2.err; echo "bench n2=$? lines=$(wc -l < gpurun_out/bench_l_n2.json)"
NCCL_DEBUG=INFO timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --alt-precisions 0 > gpurun_out/bench_l_n2_info.json 2> gpurun_out/bench_l_n2_info.err; echo "bench n2 (NCCL_DEBUG=INFO)=$? lines=$(wc -l < gpurun_out/bench_l_n2_info.json) nccl_lines=$(grep -c 'NCCL INFO' gpurun_out/bench_l_n2_info.err)"
grep -E "nRanks" gpurun_out/bench_l_n2_info.err | head -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_l_n2.json"))
print("N=2 pairs/s %.1f ms/step %.2f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
t = d["train"]
print("train all-BN-synced: %.1f ms (%.1f pairs/s), %d peer exchanges/step; backbone-BN-per-rank: %.1f ms" % (t["ms_per_step"], t["pairs_per_s"], t["peer_exchanges_per_step"], t["variant_backbone_bn_per_rank"]["ms_per_step"]))
print(t["collective"])
PY
timeout 300 $TR --impl reference bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | head -c 300; echo

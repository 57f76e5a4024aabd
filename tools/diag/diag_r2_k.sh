#!/usr/bin/env bash
# round-2 call K (2 GPUs): peer-memory SyncBN exchange: tests, training step N=1 / N=2 variants, bench N=2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "two_gpus" --timeout 600 > gpurun_out/pytest_k1.log 2>&1; echo "2-gpu tests=$?"; tail -6 gpurun_out/pytest_k1.log
echo "== train N=1"
timeout 300 python tools/train_step.py --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_step.py --steps 5 --warmup 2"
echo "== train N=2 peer syncbn (all BN)"
timeout 300 $TR --sync-bn 2>&1 | tail -1 | tee gpurun_out/train_k_n2_peer.json | cut -c1-1400
echo "== train N=2 NCCL syncbn (all BN, no host sync)"
DMB_B200_PEER_COMM=0 timeout 300 $TR --sync-bn 2>&1 | tail -1 | tee gpurun_out/train_k_n2_nccl.json | cut -c1-700
echo "== train N=2 peer syncbn, backbone BN local"
timeout 300 $TR --sync-bn --local-backbone-bn 2>&1 | tail -1 | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_k_n2.json 2> gpurun_out/bench_k_n2.err; echo "bench n2=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_k_n2.json"))
print("N=2 pairs/s %.1f ms/step %.2f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("train", d["train"]["ms_per_step"], d["train"]["pairs_per_s"], d["train"]["collective"])
PY

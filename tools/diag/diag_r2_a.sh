#!/usr/bin/env bash
# round-2 diagnostic A (2 GPUs): the unrun round-1 probes + where the 2-GPU training step loses its time
set -u
mkdir -p gpurun_out
{
echo "== probes"
timeout 60 tools/_build/mma_mn_probe
timeout 60 tools/_build/wgrad_tc_proto 0 1
timeout 60 tools/_build/wgrad_tc_proto 1 1
echo "== train N=1"
timeout 300 python tools/train_step.py --steps 4 --warmup 2 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_step.py --steps 4 --warmup 2"
echo "== train N=2 no syncbn"
timeout 300 $TR 2>&1 | tail -1
echo "== train N=2 syncbn"
timeout 300 $TR --sync-bn 2>&1 | tail -1
echo "== train N=2 syncbn, one bucket (no overlap)"
timeout 300 $TR --sync-bn --bucket-mb 1000 2>&1 | tail -1
echo "== train N=2 syncbn, no backbone"
timeout 300 $TR --sync-bn --no-backbone 2>&1 | tail -1
echo "== nccl info"
NCCL_DEBUG=INFO timeout 300 $TR --sync-bn --steps 1 --warmup 1 2>&1 | grep -E "NCCL INFO (Connected|Channel|NVLS|comm|Using|P2P)" | head -20
nvidia-smi topo -m | head -12
} > gpurun_out/diag_r2_a.log 2>&1
tail -40 gpurun_out/diag_r2_a.log

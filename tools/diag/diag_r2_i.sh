#!/usr/bin/env bash
# round-2 call I (1 GPU): Cmn confidence heads, correlation1d, SGA mode switch; full suite
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_parity.py -m gpu -q -x -k "cmn or correlation1d or sga" --timeout 300 > gpurun_out/pytest_i1.log 2>&1; echo "new tests=$?"; tail -15 gpurun_out/pytest_i1.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu_i.log
timeout 200 python - <<'PY'
import sys, time, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import densematchingbenchmark_b200 as P
from densematchingbenchmark_b200.modeling.stereo.cmn import ConfHead
h = ConfHead(192).cuda().eval()
c = torch.randn(2, 192, 544, 960, device="cuda")
with torch.no_grad():
    for _ in range(3): h(c)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): h(c)
    e.record(); torch.cuda.synchronize()
    ours = s.elapsed_time(e) / 10
    ref = torch.nn.Sequential(torch.nn.Sequential(torch.nn.Conv2d(192, 64, 3, 1, 1, bias=False), torch.nn.BatchNorm2d(64), torch.nn.ReLU()), torch.nn.Conv2d(64, 1, 1, bias=False)).cuda().eval()
    torch.backends.cudnn.allow_tf32 = False
    for _ in range(3): ref(c)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): ref(c)
    e.record(); torch.cuda.synchronize()
    print("ConfHead [2,192,544,960]: ours %.3f ms, torch/cuDNN fp32 (TF32 off) %.3f ms" % (ours, s.elapsed_time(e) / 10))
PY

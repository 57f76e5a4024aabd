#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lga or config4 or ganet or full_size" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/q_pytest.log
cat > /tmp/lga_t.py <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
from densematchingbenchmark_b200.ops import functional as F_
x = torch.randn(1, 192, 384, 1248, device="cuda"); gd = torch.randn(1, 75, 384, 1248, device="cuda")
for _ in range(3): F_.lga(x, gd, 2)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): F_.lga(x, gd, 2)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
nb = (2 * x.numel() + gd.numel()) * 4
print("LGA mode", os.environ.get("DMB_B200_LGA_ROT"), "ms", round(ms, 4), "GB/s", round(nb / ms / 1e6, 1))
PY
for m in 1 2 0; do DMB_B200_LGA_ROT=$m timeout 300 python /tmp/lga_t.py 2>&1 | tail -1; done

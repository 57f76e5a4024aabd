#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cat > /tmp/lga_t.py <<'PY'
import sys, os, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from densematchingbenchmark_b200.ops import functional as F_
x = torch.randn(1, 192, 384, 1248, device="cuda"); gd = torch.randn(1, 75, 384, 1248, device="cuda")
for _ in range(3): y = F_.lga(x, gd, 2)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): F_.lga(x, gd, 2)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
nb = (2 * x.numel() + gd.numel()) * 4
print("LGA mode", os.environ.get("DMB_B200_LGA_ROT"), "ms", round(ms, 4), "GB/s", round(nb / ms / 1e6, 1), "checksum", float(y.double().sum()))
PY
for m in 1 3 4 2; do DMB_B200_LGA_ROT=$m timeout 300 python /tmp/lga_t.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests -m gpu -x -q -k "lga or config4 or ganet" > gpurun_out/x_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/x_pytest.log
DMB_B200_LGA_ROT=4 timeout 600 python -m pytest tests -m gpu -x -q -k "lga or config4 or ganet" > gpurun_out/x_pytest4.log 2>&1; echo "pytest(mode 4) rc $?"; tail -3 gpurun_out/x_pytest4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc|head_gather|upsample|cat_volume" -c 300 --csv \
    --log-file gpurun_out/launches_hot_path_x.csv python tools/profile_hot_path.py auto fp16x3 1 > gpurun_out/prof_x.log 2>&1; echo "ncu launches=$?"

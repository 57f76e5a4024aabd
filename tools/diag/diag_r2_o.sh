#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cat > /tmp/lga_small.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import dmb_oracle as O
from densematchingbenchmark_b200.ops import functional as F_
g = torch.Generator().manual_seed(1)
x = torch.randn(1, 9, 12, 40, generator=g); gd = torch.randn(1, 75, 12, 40, generator=g)
got = F_.lga(x.cuda(), gd.cuda(), 2)
torch.cuda.synchronize()
print("max diff", float((got.cpu() - O.lga(x, gd)).abs().max()))
PY
timeout 300 compute-sanitizer --tool memcheck python /tmp/lga_small.py 2>&1 | grep -v "^$" | head -40

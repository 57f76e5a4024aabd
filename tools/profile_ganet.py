"""One warm-up + one profiled call of SGA and LGA at BASELINE config 4 sizes (the command wrapped by ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densematchingbenchmark_b200.ops import functional as F_  # noqa: E402

g = torch.Generator().manual_seed(0)
x = torch.randn(1, 32, 64, 128, 416, generator=g).cuda()
gd = torch.randn(1, 640, 128, 416, generator=g).cuda()
for _ in range(2):
    F_.sga(x, gd)
torch.cuda.synchronize()
del x, gd
xl = torch.randn(1, 192, 384, 1248, generator=g).cuda()
gl = torch.randn(1, 75, 384, 1248, generator=g).cuda()
for _ in range(2):
    F_.lga(xl, gl, 2)
torch.cuda.synchronize()
print("done")

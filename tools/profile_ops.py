"""One warm-up + one profiled call of the volume builders / LGA at the BASELINE config sizes (the command wrapped by
ncu --set full for profiles/r2_ncu_full_*.txt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densematchingbenchmark_b200.ops import functional as F_  # noqa: E402
from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as tc  # noqa: E402

g = torch.Generator().manual_seed(0)
l = torch.randn(1, 32, 136, 240, generator=g).cuda(); r = torch.randn(1, 32, 136, 240, generator=g).cuda()
for _ in range(2):
    tc.cat_volume_blocked(l, r, 48, 0, 1, "fp16x3")
    F_.cat_volume(l, r, 48)
l3 = torch.randn(1, 320, 136, 240, generator=g).cuda(); r3 = torch.randn(1, 320, 136, 240, generator=g).cuda()
for _ in range(2):
    F_.gwc_volume(l3, r3, 40, 48)
torch.cuda.synchronize()
del l3, r3
xl = torch.randn(1, 192, 384, 1248, generator=g).cuda()
gl = torch.randn(1, 75, 384, 1248, generator=g).cuda()
for _ in range(2):
    F_.lga(xl, gl, 2)
torch.cuda.synchronize()
print("done")

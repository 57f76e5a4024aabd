"""SGA / LGA / GWC at the BASELINE config-3/4 sizes, one call each (ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densematchingbenchmark_b200.ops import functional as F_  # noqa: E402
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "sga"):
    x = torch.randn(1, 32, 64, 128, 416, generator=g).to(dev)
    gd = torch.randn(1, 4 * 5 * 32, 128, 416, generator=g).to(dev)
    for _ in range(2):
        F_.sga(x, gd)
    torch.cuda.synchronize(); del x, gd
if which in ("all", "lga"):
    xl = torch.randn(1, 192, 384, 1248, generator=g).to(dev)
    gl = torch.randn(1, 75, 384, 1248, generator=g).to(dev)
    for _ in range(2):
        F_.lga(xl, gl, 2)
    torch.cuda.synchronize(); del xl, gl
if which in ("all", "gwc"):
    l3 = torch.randn(1, 320, 136, 240, generator=g).to(dev); r3 = torch.randn(1, 320, 136, 240, generator=g).to(dev)
    for _ in range(2):
        F_.gwc_volume(l3, r3, 40, 48)
    torch.cuda.synchronize()
print("done")

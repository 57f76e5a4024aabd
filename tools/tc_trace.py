"""Timeline of one stride-1 tcgen05 layer (32->32 at 48x136x240, split fp16): CTA 0 stamps clock64() in its MMA-issue,
epilogue and TMA-producer roles (dmb_b200_debug_set_trace); prints per-plane phase durations in cycles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densematchingbenchmark_b200 import _cabi as C  # noqa: E402
from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
conv = torch.nn.Conv3d(32, 32, 3, 1, 1, bias=False).to(dev)
x = torch.randn(1, 32, 48, 136, 240, device=dev)
xb = T.Blocked.from_ncdhw(x, True, True)
for _ in range(2):
    T.conv_tc(conv, xb, relu=True)
torch.cuda.synchronize()
buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
C.call("dmb_b200_debug_set_trace", C.ptr(buf))
T.conv_tc(conv, xb, relu=True)
torch.cuda.synchronize()
C.call("dmb_b200_debug_set_trace", None)
tr = buf.cpu().view(3, 4096)
mma = tr[0][tr[0] > 0].view(-1, 4)
epi = tr[1][tr[1] > 0].view(-1, 4)
pro = tr[2][tr[2] > 0].view(-1, 2)
t0 = int(min(mma[0, 0], epi[0, 0], pro[0, 0]))
print("planes traced: mma %d, epilogue %d, producer %d; kernel span %d cycles" % (len(mma), len(epi), len(pro), int(max(mma[-1, 3], epi[-1, 3])) - t0))
print("MMA warp   : issue kd=0,1 %.0f | next plane's barriers %.0f | issue kd=2 + commits %.0f | plane period %.0f (mean cycles)" % (
    float((mma[:, 1] - mma[:, 0]).float().mean()), float((mma[:, 2] - mma[:, 1]).float().mean()),
    float((mma[:, 3] - mma[:, 2]).float().mean()), float((mma[1:, 0] - mma[:-1, 0]).float().mean())))
print("epilogue   : wait accumulator %.0f | drain TMEM %.0f | convert+store %.0f | plane period %.0f" % (
    float((epi[:, 1] - epi[:, 0]).float().mean()), float((epi[:, 2] - epi[:, 1]).float().mean()),
    float((epi[:, 3] - epi[:, 2]).float().mean()), float((epi[1:, 0] - epi[:-1, 0]).float().mean())))
print("producer   : wait free stage %.0f | plane period %.0f" % (
    float((pro[:, 1] - pro[:, 0]).float().mean()), float((pro[1:, 0] - pro[:-1, 0]).float().mean())))
print("first 16 planes, cycles since start:")
for i in range(min(16, len(mma))):
    print("  plane %2d mma %s   epi %s" % (i, [int(v) - t0 for v in mma[i]], [int(v) - t0 for v in epi[i]] if i < len(epi) else None))

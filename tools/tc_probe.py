"""Bring-up probe for the tcgen05 conv kernel: error statistics against an fp64 reference."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as tc  # noqa: E402
from densematchingbenchmark_b200.ops import functional as F_  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(1)
cin, cout, dims = 32, 32, (8, 20, 24)
x = torch.randn(1, cin, *dims, generator=g)
w = torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5
ref = F.conv3d(x.double(), w.double(), padding=1)
conv = torch.nn.Conv3d(cin, cout, 3, 1, 1, bias=False)
conv.weight.data.copy_(w)
conv = conv.to(DEV)
print("absmax ref %.3f" % float(ref.abs().max()))
f32cpu = F.conv3d(x, w, padding=1).double()
print("cpu fp32       : max %.2e  mean|e| %.2e" % (float((f32cpu - ref).abs().max()), float((f32cpu - ref).abs().mean())))
d = F_.conv3d_fused(x.to(DEV), F_.pack_conv_weight(w).to(DEV), None, (3, 3, 3), 1, 1).cpu().double()
print("direct fp32    : max %.2e  mean|e| %.2e" % (float((d - ref).abs().max()), float((d - ref).abs().mean())))
for prec in ("fp16x3", "bf16x3", "fp16", "bf16"):
    split, fp16 = tc.PRECISIONS[prec]
    xb = tc.Blocked.from_ncdhw(x.to(DEV), split, fp16)
    y = tc.conv_tc(conv, xb)
    got = y.to_ncdhw().cpu().double()
    e = got - ref
    # toward-zero bias shows up as sign(ref) * e < 0 on average
    bias = float((e * torch.sign(ref)).mean())
    print("tc %-7s     : max %.2e  mean|e| %.2e  mean(sign(ref)*e) %.2e" % (prec, float(e.abs().max()), float(e.abs().mean()), bias))
# classifier-head path (fp32 output, no 16-bit output rounding)
w1 = torch.randn(1, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5
ref1 = F.conv3d(x.double(), w1.double(), padding=1)
c1 = torch.nn.Conv3d(cin, 1, 3, 1, 1, bias=False); c1.weight.data.copy_(w1); c1 = c1.to(DEV)
for prec in ("fp16x3", "bf16x3"):
    split, fp16 = tc.PRECISIONS[prec]
    xb = tc.Blocked.from_ncdhw(x.to(DEV), split, fp16)
    got = tc.conv_tc(c1, xb).cpu().double()
    e = got - ref1
    print("tc %-7s f32out: max %.2e  mean|e| %.2e  mean(sign(ref)*e) %.2e" % (prec, float(e.abs().max()), float(e.abs().mean()), float((e * torch.sign(ref1)).mean())))

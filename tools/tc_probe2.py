"""Diagnostic for the strided tcgen05 kernels: per output-parity-class error."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as tc  # noqa: E402
DEV = "cuda:0"
for kind, cin, cout, dims in (("s2", 32, 32, (4, 32, 16)), ("tr", 32, 32, (3, 16, 8)), ("tr", 32, 32, (4, 19, 11))):
    for prec in ("fp16x3", "fp16"):
        split, fp16 = tc.PRECISIONS[prec]
        g = torch.Generator().manual_seed(3)
        x = torch.randn(1, cin, *dims, generator=g)
        conv = torch.nn.Conv3d(cin, cout, 3, 2, 1, bias=False) if kind == "s2" else torch.nn.ConvTranspose3d(cin, cout, 3, 2, 1, output_padding=1, bias=False)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.05)
        w = conv.weight.detach().clone()
        ref = F.conv3d(x, w, None, stride=2, padding=1) if kind == "s2" else F.conv_transpose3d(x, w, None, stride=2, padding=1, output_padding=1)
        got = tc.conv_tc(conv.to(DEV), tc.Blocked.from_ncdhw(x.to(DEV), split, fp16)).to_ncdhw().cpu()
        e = (got - ref).abs()
        print(kind, prec, dims, "max err %.3e ref absmax %.3f" % (float(e.max()), float(ref.abs().max())))
        if kind == "tr":
            for rd in range(2):
                for rh in range(2):
                    for rw in range(2):
                        print("   class (%d,%d,%d): max err %.3e" % (rd, rh, rw, float(e[:, :, rd::2, rh::2, rw::2].max())), end=";")
                print()
            print("   per-depth max err:", [round(float(e[:, :, d].max()), 3) for d in range(e.shape[2])])

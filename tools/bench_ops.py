"""CUDA-event timings of the individual hot-path ops at the BASELINE.json config sizes (configs 2-4),
with their algorithmic bytes and the fraction of the measured HBM copy peak.  Output: one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import densematchingbenchmark_b200 as P  # noqa: E402
from densematchingbenchmark_b200.ops import functional as F_  # noqa: E402
from densematchingbenchmark_b200.ops import GateRecurrent2dnoind  # noqa: E402
from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as tc  # noqa: E402

dev = torch.device("cuda", 0)
peak = bench.peaks()["hbm_gbs"]
out = {}


def timeit(name, fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    gbs = nbytes / (ms * 1e-3) / 1e9
    out[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB_per_s": round(gbs, 1),
                 "frac_of_hbm_peak": round(gbs / peak, 3)}
    print(name, out[name], flush=True)


g = torch.Generator().manual_seed(0)
# config 2: PSMNet cat volume, features [1,32,136,240], D4=48
l = torch.randn(1, 32, 136, 240, generator=g).to(dev); r = torch.randn(1, 32, 136, 240, generator=g).to(dev)
vol_elems = 64 * 48 * 136 * 240
timeit("cat_fms fp32 NCDHW (cfg2)", lambda: F_.cat_volume(l, r, 48), 2 * l.numel() * 4 + vol_elems * 4)
timeit("dif_fms fp32 NCDHW (cfg2 size)", lambda: F_.dif_volume(l, r, 48), 2 * l.numel() * 4 + vol_elems // 2 * 4)
if tc.tc_available():
    timeit("cat volume blocked fp16 hi+lo (cfg2)", lambda: tc.cat_volume_blocked(l, r, 48, 0, 1, "fp16x3"), 2 * l.numel() * 4 + vol_elems * 4)
    timeit("cat volume blocked fp16 single plane (cfg2)", lambda: tc.cat_volume_blocked(l, r, 48, 0, 1, "fp16"), 2 * l.numel() * 4 + vol_elems * 2)
# fused regress / materialised upsample + soft-argmin (per cost)
low = torch.randn(1, 1, 48, 136, 240, generator=g).to(dev)
timeit("upsample+soft-argmin fused (cfg2, per cost)", lambda: F_.upsample_regress(low, (192, 544, 960), want_cost=False, want_disp=True), low.numel() * 4 + 544 * 960 * 4)
timeit("upsample materialised (cfg2, per cost)", lambda: F_.upsample_regress(low, (192, 544, 960), want_cost=True, want_disp=False), low.numel() * 4 + 192 * 544 * 960 * 4)
cost = torch.randn(1, 192, 544, 960, generator=g).to(dev)
timeit("soft_argmin on a materialised cost (cfg2)", lambda: F_.soft_argmin(cost), cost.numel() * 4 + 544 * 960 * 4)
timeit("local_soft_argmin r=2 (cfg2)", lambda: F_.local_soft_argmin(cost, 2), cost.numel() * 4 + 544 * 960 * 4)
del cost
# config 3: GwcNet group-wise correlation, 320 ch, 40 groups
l3 = torch.randn(1, 320, 136, 240, generator=g).to(dev); r3 = torch.randn(1, 320, 136, 240, generator=g).to(dev)
timeit("gwc volume 40 groups fp32 (cfg3)", lambda: F_.gwc_volume(l3, r3, 40, 48), 2 * l3.numel() * 4 + 40 * 48 * 136 * 240 * 4)
del l3, r3
# config 4: GANet SGA / LGA at 1248x384 (padded 1242x375)
x = torch.randn(1, 32, 64, 128, 416, generator=g).to(dev)
gd = torch.randn(1, 4 * 5 * 32, 128, 416, generator=g).to(dev)
timeit("SGA fp32 (cfg4)", lambda: F_.sga(x, gd), (2 * x.numel() + gd.numel()) * 4, reps=3)
del x, gd
xl = torch.randn(1, 192, 384, 1248, generator=g).to(dev)
gl = torch.randn(1, 75, 384, 1248, generator=g).to(dev)
timeit("LGA r=2 fp32 (cfg4)", lambda: F_.lga(xl, gl, 2), (2 * xl.numel() + gl.numel()) * 4, reps=3)
del xl, gl
# SPN (AnyNet: C=8 at 1/4 resolution)
X = torch.randn(1, 8, 136, 240, generator=g).to(dev)
G = [torch.rand(1, 8, 136, 240, generator=g).to(dev) * 0.3 for _ in range(3)]
for hz, rev, nm in ((True, False, "left->right"), (False, False, "top->bottom")):
    mod = GateRecurrent2dnoind(hz, rev)
    timeit("SPN scan %s (8x136x240)" % nm, lambda: mod(X, *G), 5 * X.numel() * 4)
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r1_ops.json"), "w"), indent=1)

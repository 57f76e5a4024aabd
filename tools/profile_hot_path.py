"""One warm-up + N profiled passes of the PSMNet hot path at BASELINE config 2 size (features
[1,32,136,240], D=192) -- the command wrapped by ncu for the launch lists under profiles/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import seeded  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "auto"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
_, proc, pred, _ = bench.build_model(dev, engine, precision)
l, r = seeded.feature_pair(1, 32, bench.H4, bench.W4, seed=5, scale=0.5, shift=6)
l, r = l.to(dev), r.to(dev)
for i in range(1 + passes):
    with torch.no_grad():
        costs = proc(l, r)
        disps = [pred(c) for c in costs]
    torch.cuda.synchronize()
print("done", float(disps[0].mean()))

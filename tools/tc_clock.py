"""Effective SM clock under the trunk's tensor-core load: the same 32->32 layer launched back to back; kernel time
from CUDA events against the clock64() span CTA 0 records (dmb_b200_debug_set_trace), plus nvidia-smi's view."""
import os
import subprocess
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
buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
for reps in (1, 20, 200, 1000):
    T.conv_tc(conv, xb, relu=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        T.conv_tc(conv, xb, relu=True)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / reps
    smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,power.limit,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    buf.zero_()
    C.call("dmb_b200_debug_set_trace", C.ptr(buf))
    T.conv_tc(conv, xb, relu=True)
    torch.cuda.synchronize()
    C.call("dmb_b200_debug_set_trace", None)
    tr = buf.cpu().view(3, 4096)
    st = tr[tr > 0]
    span = int(st.max() - st.min())
    print("%4d back-to-back launches: %.1f us per launch; CTA 0 spans %d cycles => effective SM clock %.2f GHz (if CTA 0 spans the kernel); nvidia-smi after: %s"
          % (reps, us, span, span / us / 1e3, smi))

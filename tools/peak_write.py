"""What can a WRITE-ONLY kernel reach on this box?  The roofline denominator (MEASURED_PEAKS.json: hbm_gbs) is a
copy: half of its bytes are reads.  The cost-volume builders (cat / gwc volumes, upsampling) write 50-100x more than
they read, so the relevant ceiling is the device's fill bandwidth.  Prints one JSON line: torch fill_ (vectorised
elementwise kernel), cudaMemsetAsync (zero_) and a device-to-device copy on the same 401 MB buffers."""
import json

import torch

dev = torch.device("cuda", 0)
n = 64 * 48 * 136 * 240            # the config-2 cat volume, fp32
a = torch.empty(n, device=dev, dtype=torch.float32)
b = torch.empty(n, device=dev, dtype=torch.float32)


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


nbytes = n * 4
out = {}
for name, fn, moved in (("fill_", lambda: a.fill_(1.5), nbytes), ("zero_ (memset)", lambda: a.zero_(), nbytes),
                        ("copy_", lambda: b.copy_(a), 2 * nbytes)):
    ms = t(fn)
    out[name] = {"ms": round(ms, 4), "GB_per_s": round(moved / ms / 1e6, 1)}
print(json.dumps(out))

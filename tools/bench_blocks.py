"""Extra blocks of the bench.py JSON line (round-2 review items): the individual ops at the BASELINE config 3 / 4
sizes and the reference's own PyTorch path on the same GPU.  Imported by bench.py; every timing is CUDA events on
the current stream after warm-up."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def ops_block(device, hbm_gbs, reps=10):
    """GWC volume (config 3), SGA and LGA (config 4) and the cat volume (config 2) at full size: ms, algorithmic
    bytes (at the fp32 element size the kernels move) and fraction of the measured HBM copy
    peak.  Inputs exceed the 126 MB L2 in every case, so back-to-back repetitions do not hit in cache."""
    from densematchingbenchmark_b200.ops import functional as F_
    g = torch.Generator().manual_seed(0)
    out = {}

    def rec(name, fn, nbytes, n=reps):
        ms = _time(fn, n)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB_per_s": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / hbm_gbs, 3)}

    l = torch.randn(1, 32, 136, 240, generator=g).to(device); r = torch.randn(1, 32, 136, 240, generator=g).to(device)
    rec("cat_fms fp32 [1,64,48,136,240] (cfg 2)", lambda: F_.cat_volume(l, r, 48), 2 * l.numel() * 4 + 64 * 48 * 136 * 240 * 4)
    l3 = torch.randn(1, 320, 136, 240, generator=g).to(device); r3 = torch.randn(1, 320, 136, 240, generator=g).to(device)
    rec("gwc volume, 40 groups of 8 channels, fp32 [1,40,48,136,240] (cfg 3)", lambda: F_.gwc_volume(l3, r3, 40, 48),
        2 * l3.numel() * 4 + 40 * 48 * 136 * 240 * 4)
    del l3, r3
    x = torch.randn(1, 32, 64, 128, 416, generator=g).to(device)
    gd = torch.randn(1, 4 * 5 * 32, 128, 416, generator=g).to(device)
    rec("SGA fp32 [1,32,64,128,416] (cfg 4)", lambda: F_.sga(x, gd), (2 * x.numel() + gd.numel()) * 4, max(3, reps // 2))
    del x, gd
    xl = torch.randn(1, 192, 384, 1248, generator=g).to(device)
    gl = torch.randn(1, 75, 384, 1248, generator=g).to(device)
    rec("LGA r=2 fp32 [1,192,384,1248] (cfg 4)", lambda: F_.lga(xl, gl, 2), (2 * xl.numel() + gl.numel()) * 4, max(3, reps // 2))
    del xl, gl
    torch.cuda.empty_cache()
    return out


def gpu_torch_baseline_block(sd, lf, rf, max_disp, ours_hot_path_ms, reps=5):
    """the reference's own PyTorch arithmetic for the hot path (cat volume by slice
    assignment, cuDNN 3-D convolutions + BatchNorm + ReLU + adds as separate ops, F.interpolate, softmax, expectation:
    the oracle port run on the same GPU) in fp32 with TF32 disabled and under bf16 autocast, beside our hot path on
    the same features.  Favourable to the reference in one respect: its cat_fms allocates the volume on the CPU and
    uploads 401 MB per call (cat_fms.py:32); here it is allocated on the device."""
    import dmb_oracle as O
    B = lf.shape[0]
    sd_dev = {k: v.to(lf.device) for k, v in sd.items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    try:
        with torch.no_grad():
            def fp32():
                return O.psm_hot_path(sd_dev, lf, rf, max_disp, prefix="")[1]

            def bf16():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.psm_hot_path(sd_dev, lf, rf, max_disp, prefix="")[1]

            ms32 = _time(fp32, reps, warm=2)
            d32 = fp32()[0].float()
            ms16 = _time(bf16, reps, warm=2)
            d16 = bf16()[0].float()
        res = {
            "what": "oracle port of the reference hot path on the same GPU (torch/cuDNN), %d pair(s) per call" % B,
            "fp32_tf32_off": {"hot_path_ms": ms32, "pairs_per_s": B / (ms32 * 1e-3)},
            "bf16_autocast": {"hot_path_ms": ms16, "pairs_per_s": B / (ms16 * 1e-3),
                              "max_abs_disp_diff_vs_fp32_px": float((d16 - d32).abs().max())},
            "ours_hot_path_ms": ours_hot_path_ms,
            "speedup_vs_fp32": ms32 / ours_hot_path_ms, "speedup_vs_bf16_autocast": ms16 / ours_hot_path_ms,
        }
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        del sd_dev
        torch.cuda.empty_cache()
    return res

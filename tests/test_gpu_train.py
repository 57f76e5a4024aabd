"""GPU parity of the TRAINING path (BASELINE config 5): every backward kernel against torch autograd of the
CPU oracle, the whole training step against the fixture produced by the reference's own modules
(tests/golden/train_step.pt), and the data-parallel step (NCCL gradient all-reduce + synchronised
BatchNorm) against the oracle's sharded restatement.

Tolerances: gradients are sums of up to ~1e5 fp32 products accumulated in a different order than ATen's,
so they are compared relative to the tensor's own scale (rtol 2e-3 of max |g|); forward values at the
fp32 rounding level."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

import dmb_oracle as O
import seeded
from make_golden import train_inputs
from test_oracle_golden import check_grad_summary

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import densematchingbenchmark_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg


def close(got, want, rtol=2e-3, what=""):
    scale = float(want.abs().max())
    err = float((got.detach().cpu() - want).abs().max())
    assert err <= rtol * scale + 1e-6, "%s: max err %.3e vs scale %.3e" % (what, err, scale)


# ---------------------------------------------------------------------------- conv + bn (+relu) units
UNIT_CASES = [
    # name, cin, cout, stride, transposed, bias, bn, relu, residual, dims (B, D, H, W)
    ("s1_bn_relu", 32, 32, 1, False, False, True, True, False, (2, 6, 5, 37)),
    ("s1_bn_res", 32, 32, 1, False, True, True, False, True, (2, 4, 6, 40)),
    ("s1_wide_in", 64, 32, 1, False, False, True, True, False, (1, 4, 4, 72)),
    ("s2_bn_relu", 32, 64, 2, False, False, True, True, False, (2, 8, 6, 20)),
    ("s2_odd", 8, 16, 2, False, True, True, True, False, (1, 5, 7, 9)),
    ("t2_bn_res_relu", 64, 64, 2, True, False, True, False, True, (2, 3, 4, 10)),
    ("t2_bn", 64, 32, 2, True, False, True, False, False, (1, 4, 3, 18)),
    ("head_plain", 32, 1, 1, False, False, False, False, True, (2, 4, 5, 33)),
    ("nobn_bias_relu", 16, 24, 1, False, True, False, True, False, (1, 3, 4, 11)),
]


@pytest.mark.parametrize("case", UNIT_CASES, ids=[c[0] for c in UNIT_CASES])
def test_conv_unit_train_forward_backward(P, case):
    from densematchingbenchmark_b200.modeling.stereo.layers import basic_layers as L
    name, cin, cout, stride, transposed, bias, bn, relu, with_res, dims = case
    B, D, H, W = dims
    g = torch.Generator().manual_seed(len(name) + cin)
    if transposed:
        unit = L.FusedConvUnit(torch.nn.ConvTranspose3d(cin, cout, 3, stride, 1, output_padding=1, bias=bias),
                               torch.nn.BatchNorm3d(cout) if bn else None, relu=relu)
    else:
        unit = L.FusedConvUnit(torch.nn.Conv3d(cin, cout, 3, stride, 1, bias=bias),
                               torch.nn.BatchNorm3d(cout) if bn else None, relu=relu)
    with torch.no_grad():
        for p in unit.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.2 if p.dim() > 1 else 0.5) + (1.0 if p.dim() == 1 else 0.0))
        if bn:
            unit[1].running_mean.copy_(torch.randn(cout, generator=g) * 0.1)
            unit[1].running_var.copy_(torch.rand(cout, generator=g) + 0.5)
    x = torch.randn(B, cin, D, H, W, generator=g)

    # ---- torch autograd on CPU (what the reference's nn.Sequential does)
    if transposed:
        ref_conv = torch.nn.ConvTranspose3d(cin, cout, 3, stride, 1, output_padding=1, bias=bias)
    else:
        ref_conv = torch.nn.Conv3d(cin, cout, 3, stride, 1, bias=bias)
    ref = torch.nn.Sequential(*([ref_conv] + ([torch.nn.BatchNorm3d(cout)] if bn else []))).train()
    ref.load_state_dict({k: v.clone() for k, v in unit.state_dict().items()})
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    res = torch.randn(yr.shape, generator=g) if with_res else None
    rr = res.clone().requires_grad_(True) if with_res else None
    if with_res:
        yr = yr + rr
    relu_after = with_res and name.endswith("relu")
    if relu or relu_after:
        yr = F.relu(yr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)

    # ---- ours
    unit = unit.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    rg = res.to(DEV).requires_grad_(True) if with_res else None
    y = unit(xg, residual=rg, relu_after=relu_after)
    y.backward(gy.to(DEV))
    close(y, yr.detach(), 1e-4, "y")
    close(xg.grad, xr.grad, what="dx")
    if with_res:
        close(rg.grad, rr.grad, 1e-5, "dres")
    got = dict(unit.named_parameters())
    for k, p in ref.named_parameters():
        if k == "0.bias" and bn:
            assert float(got[k].grad.abs().max()) == 0.0 and float(p.grad.abs().max()) < 1e-3
            continue
        close(got[k].grad, p.grad, what="d" + k)
    if bn:
        for k in ("running_mean", "running_var"):
            torch.testing.assert_close(getattr(unit[1], k).cpu(), getattr(ref[1], k), rtol=1e-4, atol=1e-5)
        assert int(unit[1].num_batches_tracked) == int(ref[1].num_batches_tracked) == 1


def test_conv_wgrad_multi_tile_and_segments(P):
    """Raw C-ABI call: 64->64 channels (4 channel tiles), W = 100 (4 row segments, ragged tail), B = 3."""
    from densematchingbenchmark_b200 import _cabi as C
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 64, 3, 4, 100, generator=g)
    w = torch.randn(64, 64, 3, 3, 3, generator=g).requires_grad_(True)
    y = F.conv3d(x, w, padding=1)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    dw = torch.zeros(27, 64, 64, device=DEV)
    xg, gg = x.to(DEV), gy.to(DEV)               # keep the device copies alive across the launches
    C.call("dmb_b200_conv3d_wgrad", C.ptr(xg), C.ptr(gg), C.ptr(dw), 3, 64, 64,
           C.int_array([3, 4, 100]), C.int_array([3, 4, 100]), 1, 1, C.stream(torch.device(DEV)))
    got = dw.permute(2, 1, 0).reshape(64, 64, 3, 3, 3)
    close(got, w.grad, 1e-3, "dw")
    with pytest.raises(C.DmbB200Error):          # geometry is validated
        C.call("dmb_b200_conv3d_wgrad", C.ptr(xg), C.ptr(gg), C.ptr(dw), 3, 64, 64,
               C.int_array([3, 4, 100]), C.int_array([3, 4, 99]), 1, 1, C.stream(torch.device(DEV)))


WGRAD_TC_CASES = [
    # B, Ca, Cg, D, H, W: ragged rows / columns against the 4 x 64 tile, single plane, several channel passes
    (1, 32, 32, 4, 6, 32), (2, 32, 32, 3, 9, 70), (1, 64, 32, 2, 4, 64), (1, 64, 64, 3, 5, 20), (3, 32, 64, 1, 3, 130),
    (1, 32, 32, 12, 16, 128),
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES, ids=["x".join(map(str, c)) for c in WGRAD_TC_CASES])
def test_conv_wgrad_tcgen05(P, case):
    """csrc/wgrad_tc.cu (MN-major tcgen05 operands over the blocked layout, three CTA populations by depth tap) against
    torch autograd's Conv3d weight gradient on the CPU (what the reference's backward computes).  bf16 split pairs
    carry ~16 bits per operand: agreement at the 1e-4 level of the tensor's scale; the SIMT fp32 kernel agrees with it
    at the same level."""
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T
    if not T.tc_available():
        pytest.skip("tcgen05 path unavailable on this device")
    B, Ca, Cg, D, H, W = case
    g = torch.Generator().manual_seed(B * 7 + W)
    x = torch.randn(B, Ca, D, H, W, generator=g)
    w = (torch.randn(Cg, Ca, 3, 3, 3, generator=g) * 0.05).requires_grad_(True)
    y = F.conv3d(x, w, padding=1)
    gy = torch.randn(y.shape, generator=g) * 0.01              # small gradients: the reason for bfloat16's exponent range
    y.backward(gy)
    xg, gg = x.to(DEV), gy.to(DEV)
    assert T.wgrad_tc_eligible(xg, gg, (3, 3, 3), 1, 1)
    dw = T.wgrad_tc(xg, gg)
    got = dw.permute(2, 1, 0).reshape(Cg, Ca, 3, 3, 3)
    close(got, w.grad, 2e-4, "dw tcgen05")
    # accumulation semantics of the raw entry point: a second call adds onto the first
    from densematchingbenchmark_b200 import _cabi as C
    ab, gb = T.Blocked.from_ncdhw(xg, True, False), T.Blocked.from_ncdhw(gg, True, False)
    C.call("dmb_b200_conv3d_wgrad_tc", C.ptr(ab.hi), C.ptr(ab.lo), C.ptr(gb.hi), C.ptr(gb.lo), C.ptr(dw), B, Ca, Cg, D, H, W, 1, 0,
           C.stream(torch.device(DEV)))
    close(dw.permute(2, 1, 0).reshape(Cg, Ca, 3, 3, 3), 2.0 * w.grad, 2e-4, "dw accumulated twice")
    with pytest.raises(C.DmbB200Error):
        C.call("dmb_b200_conv3d_wgrad_tc", C.ptr(ab.hi), C.ptr(ab.lo), C.ptr(gb.hi), C.ptr(gb.lo), C.ptr(dw), B, Ca, 48, D, H, W, 1, 0,
               C.stream(torch.device(DEV)))


WGRAD_TC_S2_CASES = [
    # transposed?, B, C(high-res tensor), C(low-res tensor), low-res D, H, W
    (False, 1, 32, 64, 2, 4, 32), (False, 2, 64, 64, 3, 5, 18), (True, 1, 64, 32, 2, 6, 40), (True, 2, 64, 64, 1, 3, 7),
    (False, 1, 32, 64, 6, 8, 64),
]


@pytest.mark.parametrize("case", WGRAD_TC_S2_CASES, ids=["x".join(map(str, c)) for c in WGRAD_TC_S2_CASES])
def test_conv_wgrad_tcgen05_stride2(P, case):
    """Stride-2 weight gradients on tcgen05 (W-parity-split operand): Conv3d(k3, s2, p1) -- Hourglass conv1 / conv3 --
    and ConvTranspose3d(k3, s2, p1, op1) -- conv5 / conv6 -- through the same `_wgrad` the autograd Function calls,
    against torch autograd on the CPU."""
    from densematchingbenchmark_b200.ops import autograd as A
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T
    if not T.tc_available():
        pytest.skip("tcgen05 path unavailable on this device")
    transposed, B, Chi, Clo, D, H, W = case
    g = torch.Generator().manual_seed(B * 5 + W)
    if transposed:            # input low-res (Clo channels), output high-res (Chi channels)
        conv = torch.nn.ConvTranspose3d(Clo, Chi, 3, 2, 1, output_padding=1, bias=False)
        x = torch.randn(B, Clo, D, H, W, generator=g)
    else:                     # input high-res (Chi channels), output low-res (Clo channels)
        conv = torch.nn.Conv3d(Chi, Clo, 3, 2, 1, bias=False)
        x = torch.randn(B, Chi, 2 * D, 2 * H, 2 * W, generator=g)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.05)
    y = conv(x)
    gy = torch.randn(y.shape, generator=g) * 0.01
    y.backward(gy)
    xg, gg = x.to(DEV), gy.to(DEV)
    assert A._wgrad_on_tc(xg, gg, transposed, (3, 3, 3), 2, 1)
    got = A._wgrad(xg, gg, transposed, (3, 3, 3), 2, 1, conv.weight.shape)
    close(got, conv.weight.grad, 2e-4, "dw stride 2 tcgen05")


def test_conv_wgrad_tcgen05_long_accumulation(P):
    """Config-5 sized layer (32->32 at 4 x 48 x 64 x 128 voxels: ~2000 MMAs chained into every TMEM accumulator, which
    adds with truncation) against the SIMT fp32 kernel and, on a sub-sampled set of taps, float64."""
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T
    from densematchingbenchmark_b200 import _cabi as C
    if not T.tc_available():
        pytest.skip("tcgen05 path unavailable on this device")
    B, Cc, D, H, W = 4, 32, 48, 64, 128
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(B, Cc, D, H, W, generator=g, device=DEV).relu_()        # post-ReLU activations: positive mean
    gy = torch.randn(B, Cc, D, H, W, generator=g, device=DEV) * 1e-3 + 2e-4
    dw_tc = T.wgrad_tc(x, gy)
    dw_simt = torch.zeros(27, Cc, Cc, device=DEV)
    C.call("dmb_b200_conv3d_wgrad", C.ptr(x), C.ptr(gy), C.ptr(dw_simt), B, Cc, Cc, C.int_array([D, H, W]),
           C.int_array([D, H, W]), 1, 1, C.stream(torch.device(DEV)))
    # float64 truth of the centre tap and one corner tap
    xd, gd = x.double(), gy.double()
    centre = torch.einsum("bcdhw,bkdhw->ck", xd, gd)
    corner = torch.einsum("bcdhw,bkdhw->ck", xd[:, :, :-1, :-1, :-1], gd[:, :, 1:, 1:, 1:])      # tap (0,0,0)
    for name, tap, want in (("centre", 13, centre), ("corner", 0, corner)):
        e_tc = float((dw_tc[tap].double() - want).abs().max() / want.abs().max())
        e_simt = float((dw_simt[tap].double() - want).abs().max() / want.abs().max())
        print("wgrad %s tap: tcgen05 rel err %.2e, SIMT fp32 rel err %.2e" % (name, e_tc, e_simt))
        assert e_tc < 5e-4
    close(dw_tc, dw_simt.cpu(), 1e-3, "tcgen05 vs SIMT")


# ---------------------------------------------------------------------------- upsampling / regression / volume
@pytest.mark.parametrize("shape", [(2, 3, 5, 7, 12, 20, 28), (1, 6, 4, 9, 24, 16, 36), (1, 2, 2, 3, 5, 7, 11)])
def test_upsample_trilinear_backward(P, shape):
    from densematchingbenchmark_b200.ops.autograd import UpsampleTrilinearFn
    B, Dl, Hl, Wl, D, H, W = shape
    g = torch.Generator().manual_seed(D)
    low = torch.randn(B, 1, Dl, Hl, Wl, generator=g)
    gy = torch.randn(B, D, H, W, generator=g)
    lr = low.clone().requires_grad_(True)
    F.interpolate(lr, [D, H, W], mode="trilinear", align_corners=True).squeeze(1).backward(gy)
    lg = low.to(DEV).requires_grad_(True)
    out = UpsampleTrilinearFn.apply(lg, (D, H, W))
    out.backward(gy.to(DEV))
    close(out, F.interpolate(low, [D, H, W], mode="trilinear", align_corners=True).squeeze(1), 1e-5, "cost")
    close(lg.grad, lr.grad, 1e-4, "dlow")


def test_upsample_deconv_backward(P):
    from densematchingbenchmark_b200.ops.autograd import UpsampleDeconvFn
    g = torch.Generator().manual_seed(8)
    low = torch.randn(2, 1, 3, 5, 6, generator=g)
    w = torch.randn(1, 1, 8, 8, 8, generator=g) * 0.1
    gy = torch.randn(2, 12, 20, 24, generator=g)
    lr, wr = low.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv_transpose3d(lr, wr, None, stride=4, padding=2).squeeze(1).backward(gy)
    lg, wg = low.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    UpsampleDeconvFn.apply(lg, wg, (12, 20, 24)).backward(gy.to(DEV))
    close(lg.grad, lr.grad, 1e-4, "dlow")
    close(wg.grad, wr.grad, 1e-4, "dw")


@pytest.mark.parametrize("alpha,normalize", [(1.0, True), (-2.5, True), (0.7, False)])
def test_soft_argmin_backward(P, alpha, normalize):
    pred = P.PREDICTORS["FASTER"](max_disp=24, alpha=alpha, normalize=normalize).to(DEV)
    g = torch.Generator().manual_seed(3)
    cost = torch.randn(2, 24, 9, 13, generator=g) * 3
    gy = torch.randn(2, 1, 9, 13, generator=g)
    cr = cost.clone().requires_grad_(True)
    O.soft_argmin(cr, 24, alpha=alpha, normalize=normalize).backward(gy)
    cg = cost.to(DEV).requires_grad_(True)
    d = pred(cg)
    d.backward(gy.to(DEV))
    close(d, O.soft_argmin(cost, 24, alpha=alpha, normalize=normalize), 1e-5, "disp")
    close(cg.grad, cr.grad, 1e-4, "dcost")


@pytest.mark.parametrize("shape", [(2, 4, 5, 16, 6, 0, 1), (1, 3, 4, 20, 9, -3, 2), (1, 2, 2, 6, 9, -4, 1),
                                   (1, 32, 8, 130, 40, 0, 1)])
def test_cat_volume_backward(P, shape):
    B, C, H, W, md, sd, dil = shape
    l, r = seeded.feature_pair(B, C, H, W, seed=W)
    kw = dict(max_disp=md, start_disp=sd, dilation=dil)
    lr, rr = l.clone().requires_grad_(True), r.clone().requires_grad_(True)
    vol = O.cat_volume(lr, rr, **kw)
    gy = torch.randn(vol.shape, generator=torch.Generator().manual_seed(1))
    vol.backward(gy)
    lg, rg = l.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True)
    out = P.CAT_FUNCS["default"](lg, rg, **kw)
    assert torch.equal(out.detach().cpu(), vol.detach())
    out.backward(gy.to(DEV))
    close(lg.grad, lr.grad, 1e-5, "dleft")
    close(rg.grad, rr.grad, 1e-5, "dright")


# ---------------------------------------------------------------------------- the whole training step
def _cfg(P, agg, feat_disp, max_disp):
    return P.ConfigDict(model=dict(
        batch_norm=True,
        cost_processor=dict(type="Concatenation",
                            cost_computation=dict(type="default", max_disp=feat_disp, start_disp=0, dilation=1),
                            cost_aggregator=dict(type=agg, max_disp=max_disp, in_planes=64)),
        disp_predictor=dict(type="FASTER", max_disp=max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True)))


def run_train_step(P, kind, sd, l, r, gt, case, device=DEV, sync=False, reducer=False):
    """Forward + backward of our cost processor + predictor in train() mode with the smooth-L1 loss of
    configs/PSMNet/scene_flow.py:55-63 (plain torch ops on the disparity maps: outside the path)."""
    cfg = _cfg(P, kind, case["feat_disp"], case["max_disp"])
    proc = P.build_cost_processor(cfg).to(device)
    pred = P.build_disp_predictor(cfg).to(device)
    proc.aggregator.load_state_dict(sd)
    proc.train(); pred.train()
    red = None
    if sync:
        from densematchingbenchmark_b200.utils.dist_utils import enable_sync_batchnorm, GradReducer
        assert enable_sync_batchnorm(proc) > 20
        if reducer:
            red = GradReducer(proc.parameters(), bucket_mb=1.0)
    lg, rg = l.to(device).requires_grad_(True), r.to(device).requires_grad_(True)
    costs = proc(lg, rg)
    disps = [pred(c) for c in costs]
    loss = O.disp_smooth_l1(disps, gt.to(device), case["max_disp"])
    loss.backward()
    if red is not None:
        red.finish()
    return proc, loss.detach(), [d.detach() for d in disps], lg.grad, rg.grad, red


@pytest.mark.parametrize("kind", ["PSMNet", "AcfNet"])
def test_train_step_vs_reference_golden_and_oracle(P, golden_dir, kind):
    rec = torch.load(os.path.join(golden_dir, "train_step.pt"), weights_only=False)[kind]
    c = rec["case"]
    sd = seeded.seeded_state_dict(seeded.aggregator_entries(kind, 64), seed=c["seed"])
    l, r, gt = train_inputs()
    proc, loss, disps, dl, dr, _ = run_train_step(P, kind, sd, l, r, gt, c)
    # (1) against the reference's own modules (fixture)
    assert abs(float(loss) - rec["loss"]) < 1e-4 * abs(rec["loss"])
    for a, b in zip(disps, rec["disps"]):
        assert float((a.cpu() - b).abs().max()) < 1e-3                     # the north-star forward tolerance
    close(dl, rec["dleft"], what="dleft")
    close(dr, rec["dright"], what="dright")
    params = dict(proc.aggregator.named_parameters())
    assert set(params) == set(rec["grads"])
    for k, want in rec["grads"].items():
        if k.endswith(".0.bias"):
            assert float(params[k].grad.abs().max()) < 1e-4
            continue
        check_grad_summary(params[k].grad, want, rtol=3e-3, what=k)
    after = proc.aggregator.state_dict()
    for k, want in rec["running"].items():
        torch.testing.assert_close(after[k].cpu().to(want.dtype), want, rtol=1e-4, atol=1e-5)
    # (2) every gradient element against the oracle's autograd
    want = O.train_step(sd, l, r, gt, c["max_disp"], kind)
    for k, g in want["grads"].items():
        if k.endswith(".0.bias"):
            continue
        close(params[k].grad, g, 3e-3, k)


def test_training_convs_run_on_tcgen05(P, monkeypatch):
    """Forward and input-gradient convolutions of a training step take the tensor-core kernels wherever the layer
    geometry allows (all 25 trunk units + 3 heads forward; every dgrad except the 1->32 heads'), and the result
    agrees with the SIMT fp32 kernels (DMB_B200_TRAIN_TC=0 behaviour) at the documented precision."""
    from densematchingbenchmark_b200.ops import autograd as A
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T
    if not T.tc_available():
        pytest.skip("tcgen05 path unavailable on this device")
    calls = []
    real = T.conv3d_ncdhw_tc

    def counting(x, w_packed, bias, stride, transposed, precision, residual=None, relu=False, scale=None, x_blocked=None):
        calls.append((tuple(w_packed.shape[1:]), stride, transposed, precision))
        return real(x, w_packed, bias, stride, transposed, precision, residual, relu, scale, x_blocked)

    monkeypatch.setattr(T, "conv3d_ncdhw_tc", counting)
    from make_golden import TRAIN_CASE
    sd = seeded.seeded_state_dict(seeded.aggregator_entries("PSMNet", 64), seed=TRAIN_CASE["seed"])
    l, r, gt = train_inputs()
    proc, loss, disps, dl, dr, _ = run_train_step(P, "PSMNet", sd, l, r, gt, TRAIN_CASE)
    fwd = [c for c in calls if c[3] == A.TRAIN_TC_FWD]
    bwd = [c for c in calls if c[3] == A.TRAIN_TC_BWD]
    assert len(fwd) == 4 + 3 * 6 + 3 * 2, len(fwd)          # dres0/1, three hourglasses, classif units + heads
    assert len(bwd) == 4 + 3 * 6 + 3, len(bwd)              # every dgrad but the three 1->32 head dgrads
    g_tc = {k: p.grad.clone() for k, p in proc.aggregator.named_parameters()}
    monkeypatch.setattr(A, "TRAIN_TC", False)
    n = len(calls)
    proc2, loss2, disps2, dl2, dr2, _ = run_train_step(P, "PSMNet", sd, l, r, gt, TRAIN_CASE)
    assert len(calls) == n
    assert abs(float(loss) - float(loss2)) < 1e-5 * abs(float(loss2))
    for a, b in zip(disps, disps2):
        assert float((a - b).abs().max()) < 1e-3
    close(dl, dl2.cpu(), 2e-3, "dleft tc vs simt")
    for k, p in proc2.aggregator.named_parameters():
        if k.endswith(".0.bias"):
            continue
        close(g_tc[k], p.grad.cpu(), 3e-3, k)


def test_train_mode_is_not_the_inference_engine(P):
    """train() must never route through the BN-folded tensor-core trunk (it uses running statistics)."""
    cfg = _cfg(P, "PSMNet", 8, 32)
    proc = P.build_cost_processor(cfg).to(DEV).train()
    raw = torch.zeros(1, 64, 8, 8, 16, device=DEV)
    assert proc.aggregator._use_tc(raw) is False
    assert proc.aggregator.blocked_cat_volume(torch.zeros(1, 32, 8, 16, device=DEV), torch.zeros(1, 32, 8, 16, device=DEV),
                                              max_disp=8) is None


# ---------------------------------------------------------------------------- data parallel (2 GPUs, NCCL)
def _ddp_worker(rank, world, port, kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import densematchingbenchmark_b200 as pkg
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from make_golden import TRAIN_CASE
        sd = seeded.seeded_state_dict(seeded.aggregator_entries(kind, 64), seed=TRAIN_CASE["seed"])
        l, r, gt = train_inputs()
        sl = slice(rank, rank + 1)                      # one pair per rank
        proc, loss, disps, dl, dr, red = run_train_step(pkg, kind, sd, l[sl], r[sl], gt[sl], TRAIN_CASE,
                                                        device="cuda:%d" % rank, sync=True, reducer=True)
        # numpy arrays travel through the queue by value (torch tensors go through a file-descriptor hand-over that
        # needs the sender alive when the parent un-pickles: a race with this process's exit)
        grads = {k: p.grad.cpu().numpy() for k, p in proc.aggregator.named_parameters()}
        running = {k: v.cpu().numpy() for k, v in proc.aggregator.state_dict().items() if "running_" in k}
        q.put((rank, float(loss), grads, running, dl.cpu().numpy(), red.launched_early))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["PSMNet"])
def test_data_parallel_train_step_two_gpus(P, kind):
    """2 ranks x 1 pair with synchronised BatchNorm and the overlapped gradient all-reduce == the oracle's
    2-pair step whose loss is the mean of the per-shard losses."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from make_golden import TRAIN_CASE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29750 + (os.getpid() % 100)
    procs = [ctx.Process(target=_ddp_worker, args=(rk, 2, port, kind, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = [(r[0], r[1], {k: torch.from_numpy(v) for k, v in r[2].items()}, {k: torch.from_numpy(v) for k, v in r[3].items()},
            torch.from_numpy(r[4]), r[5]) for r in res]
    sd = seeded.seeded_state_dict(seeded.aggregator_entries(kind, 64), seed=TRAIN_CASE["seed"])
    l, r, gt = train_inputs()
    want = O.train_step(sd, l, r, gt, TRAIN_CASE["max_disp"], kind, shards=2)
    assert abs(0.5 * (res[0][1] + res[1][1]) - float(want["loss"])) < 1e-4 * float(want["loss"])
    for k, g in want["grads"].items():
        if k.endswith(".0.bias"):
            continue
        close(res[0][2][k], g, 3e-3, k)
        assert torch.equal(res[0][2][k], res[1][2][k]), k          # both ranks hold the same averaged gradient
    for k, v in want["running"].items():
        if "running_" in k:
            torch.testing.assert_close(res[0][3][k], v, rtol=1e-4, atol=1e-5)
    # per-rank feature gradients are NOT averaged (they feed each rank's own backbone): rank r holds shard r's,
    # scaled by 1/world only through the loss definition of the oracle
    close(res[0][4] * 0.5, want["dleft"][0:1], 3e-3, "dleft rank 0")
    assert res[0][5] >= 1                                          # at least one bucket reduced inside backward


# ---------------------------------------------------------------------------- peer-memory statistics exchange (2 GPUs)
class _TinyBackbone(torch.nn.Module):
    """conv-BN-ReLU-conv-BN with the backbone's `_forward` entry point (what graph_backbone_views captures)."""

    def __init__(self):
        super(_TinyBackbone, self).__init__()
        self.c1, self.b1 = torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8)
        self.c2, self.b2 = torch.nn.Conv2d(8, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8)

    def _forward(self, x):
        return self.b2(self.c2(torch.relu(self.b1(self.c1(x)))))


def test_backbone_views_replayed_from_cuda_graphs_match_eager(P, monkeypatch):
    """graph_backbone_views on one GPU (plain BatchNorm in train mode): features, parameter gradients and running
    statistics of two per-view passes equal the eager ones, replay after replay."""
    from densematchingbenchmark_b200.utils import dist_utils as DU
    from densematchingbenchmark_b200.modeling.stereo.backbones.PSMNet import PSMNetBackbone
    # cuDNN picks its algorithms per call site; with TF32 allowed two algorithms differ at the 1e-3 level
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(0)
    bb = PSMNetBackbone(3).cuda().train().to(memory_format=torch.channels_last)
    state = {k: v.clone() for k, v in bb.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(2, 3, 256, 256, generator=g).cuda().contiguous(memory_format=torch.channels_last) for _ in range(2)]
    ws = [torch.randn(2, 32, 64, 64, generator=g).cuda() for _ in range(2)]

    def two_views(fns):
        for prm in bb.parameters():
            prm.grad = None
        ys = [fns[v](xs[v]) for v in range(2)]
        sum((y * w).sum() for y, w in zip(ys, ws)).backward()
        return ([y.detach().clone() for y in ys], {n: prm.grad.detach().clone() for n, prm in bb.named_parameters()},
                {k: v.clone() for k, v in bb.state_dict().items() if "running" in k or "num_batches" in k})

    e_out, e_grad, e_stat = two_views([bb._forward, bb._forward])
    views = DU.graph_backbone_views(bb, xs[0], views=2)
    for rep in range(2):
        bb.load_state_dict(state)
        g_out, g_grad, g_stat = two_views(views)
        for a, b in zip(e_out, g_out):
            assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(a.abs().max()))
        gmax = max(float(v.abs().max()) for v in e_grad.values())
        for n in e_grad:                                      # (conv biases in front of a BatchNorm have a zero gradient:
            sc = float(e_grad[n].abs().max())                 #  rounding noise only, hence the absolute floor)
            assert float((e_grad[n] - g_grad[n]).abs().max()) <= 2e-3 * sc + 5e-5 * gmax, (n, rep)
        for k in e_stat:
            torch.testing.assert_close(g_stat[k].float(), e_stat[k].float(), rtol=1e-4, atol=1e-6)


def _peer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from densematchingbenchmark_b200.utils import dist_utils as DU
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        comm = DU.peer_comm()
        assert comm is not None, "peer-memory exchange could not be set up between two GPUs of one box"
        dev = "cuda:%d" % rank
        ok = True
        g = torch.Generator().manual_seed(3)
        base = [torch.randn(r + 5, 70, generator=g, dtype=torch.float64) for r in range(world)]      # same on every rank
        for it in range(40):                       # more exchanges than ring slots, varying sizes and modes
            n = 1 + (it * 7) % 70
            mine = base[rank][it % (rank + 5), :n].to(dev)
            want_sum = sum(b[it % (r + 5), :n] for r, b in enumerate(base))
            got = comm.exchange(mine.contiguous(), "sum").cpu()
            ok = ok and torch.equal(got, want_sum) if world == 2 else ok and torch.allclose(got, want_sum)
            got32 = comm.exchange(mine.float().contiguous(), "gather").cpu()
            ok = ok and all(torch.equal(got32[r], base[r][it % (r + 5), :n].float()) for r in range(world))
        # synchronised BatchNorm of a torch module == BatchNorm of the concatenated batch in one process
        torch.manual_seed(0)
        bn_ref = torch.nn.BatchNorm2d(6).train()
        bn = DU.convert_sync_batchnorm(torch.nn.Sequential(torch.nn.BatchNorm2d(6)))[0].to(dev).train()
        assert isinstance(bn, DU.PeerSyncBatchNorm)
        xs = [torch.randn(3, 6, 5, 7, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
        xall = torch.cat(xs, 0).requires_grad_(True)
        yall = bn_ref(xall)
        gys = [torch.randn(3, 6, 5, 7, generator=torch.Generator().manual_seed(20 + r)) for r in range(world)]
        yall.backward(torch.cat(gys, 0))
        x = xs[rank].to(dev).requires_grad_(True)
        y = bn(x)
        y.backward(gys[rank].to(dev))
        sl = slice(3 * rank, 3 * rank + 3)
        ok = ok and torch.allclose(y.detach().cpu(), yall.detach()[sl], atol=1e-5, rtol=1e-5)
        ok = ok and torch.allclose(x.grad.cpu(), xall.grad[sl], atol=1e-5, rtol=1e-4)
        ok = ok and torch.allclose(bn.running_mean.cpu(), bn_ref.running_mean, atol=1e-6, rtol=1e-5)
        ok = ok and torch.allclose(bn.running_var.cpu(), bn_ref.running_var, atol=1e-6, rtol=1e-5)
        # the same layers replayed from CUDA graphs (graph_backbone_views: device-side sequence counter) == eager
        torch.manual_seed(1)
        tiny = DU.convert_sync_batchnorm(_TinyBackbone()).to(dev).train()
        state = {k: v.clone() for k, v in tiny.state_dict().items()}
        xv = [torch.randn(2, 3, 12, 16, generator=torch.Generator().manual_seed(30 + 2 * rank + v)).to(dev) for v in range(2)]
        wv = [torch.randn(2, 8, 12, 16, generator=torch.Generator().manual_seed(40 + 2 * rank + v)).to(dev) for v in range(2)]

        def two_views(fns):
            for prm in tiny.parameters():
                prm.grad = None
            ys = [fns[v](xv[v]) for v in range(2)]
            sum((y * w).sum() for y, w in zip(ys, wv)).backward()
            return ([y.detach().clone() for y in ys], [prm.grad.detach().clone() for prm in tiny.parameters()],
                    {k: v.clone() for k, v in tiny.state_dict().items() if "running" in k or "num_batches" in k})

        e_out, e_grad, e_stat = two_views([tiny._forward, tiny._forward])
        n_before = comm.device_exchanges()
        views = DU.graph_backbone_views(tiny, xv[0], views=2)
        for rep in range(2):                                  # two replays from the same initial state
            tiny.load_state_dict(state)
            g_out, g_grad, g_stat = two_views(views)
            ok = ok and all(torch.allclose(a, b, atol=1e-5, rtol=1e-5) for a, b in zip(e_out, g_out))
            ok = ok and all(torch.allclose(a, b, atol=1e-4, rtol=1e-4) for a, b in zip(e_grad, g_grad))
            ok = ok and all(torch.allclose(e_stat[k].float(), g_stat[k].float(), atol=1e-6, rtol=1e-5) for k in e_stat)
        ok = ok and comm.device_exchanges() > n_before       # the replays exchanged statistics (device-side counter)
        q.put((rank, bool(ok), comm.device_exchanges()))
        DU.close_peer_comms()
    finally:
        dist.destroy_process_group()


def test_peer_memory_exchange_and_sync_batchnorm_two_gpus(P):
    """csrc/peer_comm.cu between two processes / two GPUs: exact sums and gathers over 80 exchanges (ring-slot reuse),
    and PeerSyncBatchNorm forward / backward / running statistics against single-process BatchNorm of the joint batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 100)
    procs = [ctx.Process(target=_peer_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] >= 80

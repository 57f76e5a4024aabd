"""Round-2 advisor findings, as tests: every op either carries a backward made of library kernels or refuses to
build a graph-less result (no silent zero gradients); frozen / eval-mode BatchNorm units stay differentiable;
launches follow the tensors' device, not the process's current one."""
import pytest
import torch
import torch.nn.functional as F

import dmb_oracle as O
import seeded

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def P():
    import densematchingbenchmark_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def close(got, want, rtol, what=""):
    scale = float(want.abs().max())
    err = float((got.detach().cpu() - want).abs().max())
    assert err <= rtol * scale + 1e-6, "%s: max err %.3e vs scale %.3e" % (what, err, scale)


@pytest.mark.parametrize("shape", [(2, 4, 5, 16, 6, 0, 1), (1, 3, 4, 20, 9, -3, 2), (1, 2, 2, 6, 9, -4, 1),
                                   (1, 32, 8, 130, 24, 0, 1)])
def test_dif_volume_backward(P, shape):
    """dif_fms (dif_fms.py:7-46) now has a backward (the StereoNet configs train through it): autograd of the
    oracle's restatement is the reference."""
    B, C, H, W, md, sd, dil = shape
    l, r = seeded.feature_pair(B, C, H, W, seed=W + 1)
    kw = dict(max_disp=md, start_disp=sd, dilation=dil)
    lr, rr = l.clone().requires_grad_(True), r.clone().requires_grad_(True)
    vol = O.dif_volume(lr, rr, **kw)
    gy = torch.randn(vol.shape, generator=torch.Generator().manual_seed(2))
    vol.backward(gy)
    lg, rg = l.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True)
    out = P.DIF_FUNCS["default"](lg, rg, **kw)
    assert out.grad_fn is not None
    assert torch.equal(out.detach().cpu(), vol.detach())
    out.backward(gy.to(DEV))
    close(lg.grad, lr.grad, 1e-5, "dleft")
    close(rg.grad, rr.grad, 1e-5, "dright")


def test_ops_without_backward_refuse_to_detach_silently(P):
    from densematchingbenchmark_b200.ops import SGA, LGA
    l, r = seeded.feature_pair(1, 8, 6, 16, seed=3)
    lg, rg = l.to(DEV).requires_grad_(True), r.to(DEV)
    for fn in (lambda: P.CAT_FUNCS["fast_mode"](lg, rg, max_disp=4),
               lambda: P.DIF_FUNCS["fast_mode"](lg, rg, max_disp=4),
               lambda: P.GWC_FUNCS["default"](lg, rg, max_disp=4, num_groups=2),
               lambda: SGA()(torch.randn(1, 2, 4, 6, 16, device=DEV, requires_grad=True), torch.randn(1, 40, 6, 16, device=DEV)),
               lambda: LGA(2)(torch.randn(1, 4, 6, 16, device=DEV, requires_grad=True), torch.randn(1, 75, 6, 16, device=DEV))):
        with pytest.raises(NotImplementedError):
            fn()
        with torch.no_grad():
            assert fn().grad_fn is None                         # inference is unaffected
    cfg = P.ConfigDict(model=dict(disp_predictor=dict(type="LOCAL", max_disp=8, radius=2, start_disp=0, dilation=1, alpha=1.0)))
    pred = P.build_disp_predictor(cfg).to(DEV)
    with pytest.raises(NotImplementedError):
        pred(torch.randn(1, 8, 6, 16, device=DEV, requires_grad=True))


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("mode", ["eval_with_input_grad", "train_with_frozen_bn"])
def test_frozen_batchnorm_unit_is_differentiable(P, transposed, mode):
    """`model.train(); bn.eval()` (the usual freeze-BN recipe) and an eval-mode unit fed by a tensor that requires
    grad: running statistics are used AND left untouched, gradients reach the input, the conv weight and the
    BatchNorm affine parameters -- like nn.Sequential(Conv3d, BatchNorm3d.eval(), ReLU) under torch autograd."""
    from densematchingbenchmark_b200.modeling.stereo.layers import basic_layers as L
    g = torch.Generator().manual_seed(7)
    cin, cout = 32, 32
    mk = (lambda: torch.nn.ConvTranspose3d(cin, cout, 3, 2, 1, output_padding=1, bias=True)) if transposed else \
         (lambda: torch.nn.Conv3d(cin, cout, 3, 1, 1, bias=True))
    unit = L.FusedConvUnit(mk(), torch.nn.BatchNorm3d(cout), relu=True)
    ref = torch.nn.Sequential(mk(), torch.nn.BatchNorm3d(cout), torch.nn.ReLU())
    with torch.no_grad():
        for p in unit.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.1 if p.dim() > 1 else 0.5) + (1.0 if p.dim() == 1 else 0.0))
        unit[1].running_mean.copy_(torch.randn(cout, generator=g) * 0.1)
        unit[1].running_var.copy_(torch.rand(cout, generator=g) + 0.5)
    ref.load_state_dict(unit.state_dict())
    x = torch.randn(2, cin, 4, 6, 34, generator=g)
    if mode == "train_with_frozen_bn":
        unit.train(); unit[1].eval(); ref.train(); ref[1].eval()
    else:
        unit.eval(); ref.eval()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    unit = unit.to(DEV)
    before = (unit[1].running_mean.clone(), unit[1].running_var.clone(), int(unit[1].num_batches_tracked))
    xg = x.to(DEV).requires_grad_(True)
    y = unit(xg)
    close(y, yr.detach(), 1e-5, "forward")
    y.backward(gy.to(DEV))
    close(xg.grad, xr.grad, 3e-3, "dx")
    close(unit[0].weight.grad, ref[0].weight.grad, 3e-3, "dweight")
    close(unit[0].bias.grad, ref[0].bias.grad, 3e-3, "dbias")
    close(unit[1].weight.grad, ref[1].weight.grad, 3e-3, "dgamma")
    close(unit[1].bias.grad, ref[1].bias.grad, 3e-3, "dbeta")
    assert torch.equal(unit[1].running_mean, before[0]) and torch.equal(unit[1].running_var, before[1])
    assert int(unit[1].num_batches_tracked) == before[2]


def test_eval_mode_with_grad_inputs_builds_a_graph_through_the_whole_processor(P):
    """Inference without torch.no_grad() on features that require grad (a trainable backbone in front): the
    processor must not hand back graph-less costs -- it takes the differentiable route instead of the tcgen05 engine."""
    cfg = P.ConfigDict(model=dict(
        batch_norm=True,
        cost_processor=dict(type="Concatenation", cost_computation=dict(type="default", max_disp=8, start_disp=0, dilation=1),
                            cost_aggregator=dict(type="PSMNet", max_disp=32, in_planes=64)),
        disp_predictor=dict(type="FASTER", max_disp=32, start_disp=0, dilation=1, alpha=1.0, normalize=True)))
    proc = P.build_cost_processor(cfg).to(DEV).eval()
    pred = P.build_disp_predictor(cfg).to(DEV).eval()
    l, r = seeded.feature_pair(1, 32, 8, 16, seed=5)
    with torch.no_grad():
        want = [pred(c) for c in proc(l.to(DEV), r.to(DEV))]
    lg = l.to(DEV).requires_grad_(True)
    disps = [pred(c) for c in proc(lg, r.to(DEV))]
    assert all(d.grad_fn is not None for d in disps)
    for d, w in zip(disps, want):
        assert float((d.detach() - w).abs().max()) < 1e-3
    sum(d.sum() for d in disps).backward()
    assert lg.grad is not None and float(lg.grad.abs().max()) > 0


def test_launch_follows_the_tensor_device(P):
    """Tensors on cuda:1 while cuda:0 is current (nn.DataParallel, model.to('cuda:1')): the library call must run on
    the tensors' device; mixing devices in one call is an error."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from densematchingbenchmark_b200 import _cabi
    l, r = seeded.feature_pair(1, 4, 6, 16, seed=1)
    torch.cuda.set_device(0)
    out = P.CAT_FUNCS["default"](l.to("cuda:1"), r.to("cuda:1"), max_disp=4)
    assert out.device == torch.device("cuda:1")
    assert torch.equal(out.cpu(), O.cat_volume(l, r, 4))
    with pytest.raises(_cabi.DmbB200Error):
        P.CAT_FUNCS["default"](l.to("cuda:0"), r.to("cuda:1"), max_disp=4)

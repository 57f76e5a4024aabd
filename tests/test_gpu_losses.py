"""GPU parity of the fused stereo focal loss against the fixture produced by the
reference's own StereoFocalLoss (tests/golden/focal_loss.pt) and against the CPU oracle at a larger size.
Tolerances: fp32 in a different summation order, and (1 - q)^(-5) amplifies the rounding of q near a peaked
ground-truth distribution; loss 1e-4 relative, gradients 3e-4 of the tensor's scale."""
import os

import pytest
import torch

import dmb_oracle as O
from make_golden import FOCAL_CASES, focal_inputs

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def P():
    import densematchingbenchmark_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg


def close(got, want, rtol, what):
    scale = max(float(want.abs().max()), 1e-12)
    err = float((got.detach().cpu() - want).abs().max())
    assert err <= rtol * scale + 1e-9, "%s: max err %.3e vs scale %.3e" % (what, err, scale)


@pytest.mark.parametrize("case", FOCAL_CASES, ids=[c[0] for c in FOCAL_CASES])
def test_focal_loss_vs_reference_golden(P, golden_dir, case):
    name, B, D, H, W, Hg, Wg, max_disp, start, dil, fc, sparse, vkind = case
    want = torch.load(os.path.join(golden_dir, "focal_loss.pt"), weights_only=False)[name]
    cost, gt, var = focal_inputs(case)
    cost = cost.to(DEV).requires_grad_(True)
    if torch.is_tensor(var):
        var = var.to(DEV).requires_grad_(True)
    ev = P.StereoFocalLoss(max_disp=max_disp, start_disp=start, dilation=dil, weights=(0.7,), focal_coefficient=fc,
                           sparse=sparse)
    loss = ev(cost, gt.to(DEV), var)["stereo_focal_loss_lvl0"]
    assert abs(float(loss) - want["loss"]) <= 1e-4 * max(1.0, abs(want["loss"])), (float(loss), want["loss"])
    loss.backward()
    close(cost.grad, want["dcost"], 3e-4, "dcost")
    if want["dvar"] is not None:
        close(var.grad, want["dvar"], 3e-4, "dvar")


def test_focal_loss_list_api_and_per_pixel_samples(P):
    """Three levels like AcfNet (weights 1.0 / 0.7 / 0.5, per-level variance maps) and the per-pixel disp_sample
    argument, against the oracle."""
    g = torch.Generator().manual_seed(9)
    B, D, H, W = 2, 48, 24, 40
    costs = [torch.randn(B, D, H, W, generator=g) * 4 for _ in range(3)]
    gt = torch.rand(B, 1, H, W, generator=g) * 52 - 2
    variances = [torch.rand(B, 1, H, W, generator=g) + 0.4 for _ in range(3)]
    ev = P.StereoFocalLoss(max_disp=48, weights=(1.0, 0.7, 0.5), focal_coefficient=5.0)
    cg = [c.to(DEV).requires_grad_(True) for c in costs]
    vg = [v.to(DEV).requires_grad_(True) for v in variances]
    out = ev(cg, gt.to(DEV), vg)
    assert sorted(out) == ["stereo_focal_loss_lvl0", "stereo_focal_loss_lvl1", "stereo_focal_loss_lvl2"]
    sum(out.values()).backward()
    for i, wgt in enumerate((1.0, 0.7, 0.5)):
        c = costs[i].clone().requires_grad_(True)
        v = variances[i].clone().requires_grad_(True)
        want = wgt * O.stereo_focal_loss(c, gt, v, 48, focal_coefficient=5.0)
        want.backward()
        assert abs(float(out["stereo_focal_loss_lvl%d" % i]) - float(want)) <= 1e-4 * abs(float(want))
        close(cg[i].grad, c.grad, 3e-4, "dcost lvl%d" % i)
        close(vg[i].grad, v.grad, 3e-4, "dvar lvl%d" % i)
    # per-pixel samples: the default linspace handed in explicitly must give the same loss
    ds = torch.linspace(0, 47, 48).view(1, 48, 1, 1).expand(B, 48, H, W).contiguous()
    one = P.StereoFocalLoss(max_disp=48, focal_coefficient=5.0)
    a = one(costs[0].to(DEV), gt.to(DEV), 1.3)["stereo_focal_loss_lvl0"]
    b = one(costs[0].to(DEV), gt.to(DEV), 1.3, disp_sample=ds.to(DEV))["stereo_focal_loss_lvl0"]
    assert abs(float(a) - float(b)) <= 1e-6 * abs(float(a))
    want = O.stereo_focal_loss(costs[0], gt, 1.3, 48, focal_coefficient=5.0)
    assert abs(float(a) - float(want)) <= 1e-4 * abs(float(want))


def test_focal_loss_rejects_cpu_tensors(P):
    ev = P.StereoFocalLoss(max_disp=8)
    with pytest.raises(Exception):
        ev(torch.zeros(1, 8, 4, 4), torch.ones(1, 1, 4, 4), 1.0)


# ---------------------------------------------------------------------------- confidence heads
def _cmn_cfg(P):
    from make_golden import CMN_CASE as c
    return P.ConfigDict(model=dict(batch_norm=True, cmn=dict(
        in_planes=c["in_planes"], num=c["num"], alpha=c["alpha"], beta=c["beta"],
        losses=dict(nll_loss=dict(max_disp=192, weights=(1.0, 0.7), weight=8.0)))), data=dict(sparse=False))


def test_cmn_eval_vs_reference_golden(P, golden_dir):
    """Cmn.forward in eval mode (cmn/cmn.py:57-82): confidences and confidence-modulated variances of the reference's
    own module on seeded cost volumes; ours = flat tcgen05 conv (BatchNorm folded) + blocked dot."""
    from make_golden import cmn_inputs
    from densematchingbenchmark_b200.modeling.stereo.cmn import build_cmn
    rec = torch.load(os.path.join(golden_dir, "cmn.pt"), weights_only=False)["eval"]
    costs, gt, sd = cmn_inputs()
    m = build_cmn(_cmn_cfg(P))
    m.load_state_dict(sd)                                   # the reference's key layout loads unchanged
    m = m.to(DEV).eval()
    with torch.no_grad():
        cost_vars, confs = m([t.to(DEV) for t in costs], gt.to(DEV))
    for got, want in zip(confs, rec["confs"]):
        assert float((got.cpu() - want).abs().max()) < 2e-5
    for got, want in zip(cost_vars, rec["cost_vars"]):
        assert float((got.cpu() - want).abs().max()) < 2e-5


def test_cmn_train_vs_reference_golden(P, golden_dir):
    """Training mode: batch statistics, NLL confidence losses, gradients w.r.t. the cost volumes and every parameter,
    running-statistic updates -- against the reference module's autograd (fixture summaries)."""
    from make_golden import cmn_inputs
    from test_oracle_golden import check_grad_summary
    from densematchingbenchmark_b200.modeling.stereo.cmn import build_cmn
    rec = torch.load(os.path.join(golden_dir, "cmn.pt"), weights_only=False)["train"]
    costs, gt, sd = cmn_inputs()
    m = build_cmn(_cmn_cfg(P))
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    xs = [t.to(DEV).requires_grad_(True) for t in costs]
    cost_vars, losses = m(xs, gt.to(DEV))
    for k, v in rec["losses"].items():
        assert abs(float(losses[k]) - v) < 2e-4 * max(1.0, abs(v)), (k, float(losses[k]), v)
    for got, want in zip(cost_vars, rec["cost_vars"]):
        assert float((got.detach().cpu() - want).abs().max()) < 1e-4
    (sum(losses.values()) + sum(v.mean() for v in cost_vars)).backward()
    for x, want in zip(xs, rec["dcost"]):
        check_grad_summary(x.grad.cpu(), want, 3e-3)
    for k, p in m.named_parameters():
        check_grad_summary(p.grad.cpu(), rec["grads"][k], 3e-3)
    state = m.state_dict()
    for k, v in rec["running"].items():
        torch.testing.assert_close(state[k].cpu().to(v.dtype), v, rtol=1e-4, atol=1e-5)

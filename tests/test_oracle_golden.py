"""Pins the CPU oracle (oracle/dmb_oracle.py) against outputs of the reference itself
(tests/golden/*.pt, produced by oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

import dmb_oracle as O
import seeded
from make_golden import VOLUME_CASES, PRED_CASES, volume_inputs


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("case", [c[0] for c in VOLUME_CASES])
def test_volumes_match_reference(golden_dir, case):
    rec = _load(golden_dir, "volumes.pt")[case]
    B, C, H, W, md, sd, dil = rec["params"]
    l, r = volume_inputs(case, B, C, H, W)
    kw = dict(max_disp=md, start_disp=sd, dilation=dil)
    assert torch.equal(O.cat_volume(l, r, **kw), rec["cat"])          # pure copy: bit exact
    assert torch.equal(O.dif_volume(l, r, **kw), rec["dif"])          # one fp32 subtract: bit exact
    if "fast_cat" in rec:
        torch.testing.assert_close(O.fast_cat_volume(l, r, **kw), rec["fast_cat"], atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(O.fast_dif_volume(l, r, **kw), rec["fast_dif"], atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(O.fast_dif_volume(l, r, normalize=True, p=1.0, **kw), rec["fast_dif_norm"],
                                   atol=1e-5, rtol=1e-5)
        got = O.fast_cat_volume(l, r, disp_sample=rec["disp_sample"], **kw)
        torch.testing.assert_close(got, rec["fast_cat_sampled"], atol=1e-5, rtol=1e-5)


def test_reference_test_vector_by_hand(golden_dir):
    """The reference's own parameterisation (test_cat_fms.py:26-40): disparities -2, 0, 2."""
    assert O.disp_indices(5, -2, 2) == [-2, 0, 2]
    rec = _load(golden_dir, "volumes.pt")["ref_test"]
    cat = rec["cat"]
    l, r = volume_inputs("ref_test", 1, 1, 3, 4)
    # d=+2: columns 2..3 hold L[x], R[x-2]
    assert torch.equal(cat[0, 0, 2, :, 2:], l[0, 0, :, 2:]) and torch.equal(cat[0, 1, 2, :, 2:], r[0, 0, :, :2])
    assert float(cat[0, :, 2, :, :2].abs().sum()) == 0.0
    # d=-2: columns 0..1 hold L[x], R[x+2]
    assert torch.equal(cat[0, 1, 0, :, :2], r[0, 0, :, 2:])


@pytest.mark.parametrize("case", [c[0] for c in PRED_CASES])
def test_predictors_match_reference(golden_dir, case):
    rec = _load(golden_dir, "predictors.pt")[case]
    B, md, H, W, sd, dil, alpha, norm = rec["params"]
    cost = rec["cost"]
    kw = dict(max_disp=md, start_disp=sd, dilation=dil, alpha=alpha, normalize=norm)
    got = O.soft_argmin(cost, **kw)
    torch.testing.assert_close(got, rec["DEFAULT"], atol=2e-5, rtol=1e-5)
    torch.testing.assert_close(got, rec["FASTER"], atol=2e-5, rtol=1e-5)
    torch.testing.assert_close(O.soft_argmin(cost, disp_sample=rec["disp_sample"], **kw), rec["DEFAULT_sampled"],
                               atol=2e-5, rtol=1e-5)
    for radius, rdil in ((1, 1), (2, 1), (2, 2)):
        got = O.local_soft_argmin(cost, md, radius, sd, dil, rdil, alpha)
        torch.testing.assert_close(got, rec["LOCAL_r%d_d%d" % (radius, rdil)], atol=2e-5, rtol=1e-5)


def test_uniform_cost_regresses_to_zero(golden_dir):
    """Known answer implied by test_disp_predictors.py:42-77: all-ones cost, samples -4..4."""
    rec = _load(golden_dir, "predictors.pt")["ref_test_ones"]
    assert float(rec["DEFAULT"].abs().max()) < 1e-6
    assert float(O.soft_argmin(rec["cost"], 9, -4, 2).abs().max()) < 1e-6


def test_hourglass_matches_reference(golden_dir):
    rec = _load(golden_dir, "hourglass.pt")
    entries = []
    seeded._hourglass(entries, "hg", 32, bias=False)
    sd = seeded.seeded_state_dict([(k[3:], s, r) for k, s, r in entries], seed=7)
    assert abs(seeded.checksum(sd) - rec["weight_checksum"]) < 1e-6 * rec["weight_checksum"]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 32, 8, 8, 12, generator=g)
    pre = torch.randn(1, 64, 4, 4, 6, generator=g)
    post = torch.randn(1, 64, 4, 4, 6, generator=g)
    sd = {"hg." + k: v for k, v in sd.items()}
    for got, want in zip(O.hourglass(sd, "hg", x), rec["first"]):
        torch.testing.assert_close(got, want, atol=1e-4, rtol=1e-4)
    for got, want in zip(O.hourglass(sd, "hg", x, pre, post), rec["second"]):
        torch.testing.assert_close(got, want, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("agg", ["PSMNet", "AcfNet"])
@pytest.mark.parametrize("variant", ["plain", "sharp"])
def test_aggregators_match_reference(golden_dir, agg, variant):
    rec = _load(golden_dir, "aggregators.pt")["%s_%s" % (agg, variant)]
    sd = seeded.seeded_state_dict(seeded.aggregator_entries(agg, 64), seed=rec["seed"], sharpen=rec["sharpen"])
    assert abs(seeded.checksum(sd) - rec["weight_checksum"]) < 1e-6 * rec["weight_checksum"]
    l, r = seeded.feature_pair(1, 32, 16, 32, seed=100 + rec["seed"], scale=0.5, shift=rec["shift"])
    raw = O.cat_volume(l, r, 12, 0, 1)
    fn = O.psm_aggregator if agg == "PSMNet" else O.acf_aggregator
    costs = fn(sd, raw, 48)
    for c, want, amax in zip(costs, rec["cost_samples"], rec["cost_absmax"]):
        torch.testing.assert_close(c[:, ::3, ::4, ::4], want, atol=1e-5 * max(1.0, amax), rtol=1e-4)
    for c, want in zip(costs, rec["disps"]):
        got = O.soft_argmin(c, 48)
        assert float((got - want).abs().max()) < 1e-3   # the north-star tolerance, oracle vs reference


def test_trilinear_restatement_equals_interpolate():
    g = torch.Generator().manual_seed(4)
    c = torch.randn(1, 1, 5, 6, 7, generator=g)
    want = torch.nn.functional.interpolate(c, [20, 24, 28], mode="trilinear", align_corners=True)
    torch.testing.assert_close(O.trilinear_up(c, (20, 24, 28)), want, atol=1e-5, rtol=1e-5)


def test_epe_matches_reference(golden_dir):
    rec = _load(golden_dir, "epe.pt")
    assert abs(O.epe(rec["est"], rec["gt"], 0, 192) - rec["epe"]) < 1e-6


def test_spn_scan_forward_backward_consistent():
    """spn_scan (in-place restatement) vs its autograd twin, all four directions."""
    g = torch.Generator().manual_seed(9)
    X = torch.randn(2, 3, 5, 6, generator=g)
    G = [torch.rand(2, 3, 5, 6, generator=g) * 0.3 for _ in range(3)]
    for horizontal in (True, False):
        for reverse in (False, True):
            a = O.spn_scan(X, *G, horizontal, reverse)
            b = O._spn_autograd(X, *G, horizontal, reverse)
            torch.testing.assert_close(a, b, atol=1e-6, rtol=1e-6)


def test_sga_lga_basic_properties():
    g = torch.Generator().manual_seed(12)
    x = torch.randn(1, 2, 6, 5, 7, generator=g)
    # guidance with only the w0 tap non-zero => A_r = x for all r => output x
    gd = torch.zeros(1, 4, 5, 2, 5, 7); gd[:, :, 0] = 3.0
    torch.testing.assert_close(O.sga(x, gd.view(1, 40, 5, 7)), x)
    c = torch.randn(1, 6, 5, 7, generator=g)
    gl = torch.zeros(1, 3, 5, 5, 5, 7); gl[:, 0, 2, 2] = 2.0        # centre tap only
    torch.testing.assert_close(O.lga(c, gl.view(1, 75, 5, 7)), c)


def check_grad_summary(got, want, rtol=2e-3, what=""):
    """Compare a gradient tensor with the (sum, abs-sum, strided sample) summary stored in the fixture."""
    f = got.detach().reshape(-1).cpu()
    scale = want["abssum"] / max(f.numel(), 1)                  # mean |g|: the absolute scale of this tensor
    sample = f[::want["step"]][:want["sample"].numel()]
    assert float((sample - want["sample"]).abs().max()) <= rtol * float(want["sample"].abs().max()) + 20 * rtol * scale + 1e-7, what
    assert abs(float(f.double().abs().sum()) - want["abssum"]) <= rtol * want["abssum"] + 1e-6, what
    assert abs(float(f.double().sum()) - want["sum"]) <= rtol * want["abssum"] + 1e-6, what


@pytest.mark.parametrize("kind", ["PSMNet", "AcfNet"])
def test_train_step_matches_reference(golden_dir, kind):
    """Training-mode forward/backward of the oracle (batch-statistics BatchNorm, smooth-L1) against the
    reference modules' own autograd (fixture made by oracle/make_golden.py:gen_train_step)."""
    from make_golden import train_inputs
    rec = _load(golden_dir, "train_step.pt")[kind]
    c = rec["case"]
    sd = seeded.seeded_state_dict(seeded.aggregator_entries(kind, 64), seed=c["seed"])
    assert abs(seeded.checksum(sd) - rec["weight_checksum"]) < 1e-3 * rec["weight_checksum"]
    l, r, gt = train_inputs()
    got = O.train_step(sd, l, r, gt, c["max_disp"], kind)
    assert abs(float(got["loss"]) - rec["loss"]) < 1e-4 * abs(rec["loss"])
    for a, b in zip(got["disps"], rec["disps"]):
        assert float((a - b).abs().max()) < 1e-3
    torch.testing.assert_close(got["dleft"], rec["dleft"], rtol=1e-3, atol=1e-3 * float(rec["dleft"].abs().max()))
    torch.testing.assert_close(got["dright"], rec["dright"], rtol=1e-3, atol=1e-3 * float(rec["dright"].abs().max()))
    assert set(got["grads"]) == set(rec["grads"])
    for k, want in rec["grads"].items():
        if k.endswith(".0.bias"):
            # a conv bias in front of a batch norm: the true gradient is 0, autograd leaves rounding noise
            assert float(got["grads"][k].abs().max()) < 1e-4
            continue
        check_grad_summary(got["grads"][k], want, what=k)
    for k, want in rec["running"].items():
        torch.testing.assert_close(got["running"][k].to(want.dtype), want, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ stereo focal loss
def test_focal_loss_oracle_vs_reference_golden(golden_dir):
    """oracle.stereo_focal_loss (value and autograd) == the reference's StereoFocalLoss on the seeded cases of
    oracle/make_golden.py:FOCAL_CASES (resolution change, sparse pooling, focal coefficient, tensor variance,
    start / dilation, no valid pixel)."""
    from make_golden import FOCAL_CASES, focal_inputs
    rec = torch.load(os.path.join(golden_dir, "focal_loss.pt"), weights_only=False)
    for case in FOCAL_CASES:
        name, B, D, H, W, Hg, Wg, max_disp, start, dil, fc, sparse, vkind = case
        cost, gt, var = focal_inputs(case)
        cost = cost.clone().requires_grad_(True)
        if torch.is_tensor(var):
            var = var.clone().requires_grad_(True)
        loss = 0.7 * O.stereo_focal_loss(cost, gt, var, max_disp, start, dil, fc, sparse)
        if loss.requires_grad:
            loss.backward()
        want = rec[name]
        assert abs(float(loss) - want["loss"]) <= 1e-6 * max(1.0, abs(want["loss"])), name
        got_dc = cost.grad if cost.grad is not None else torch.zeros_like(cost)
        torch.testing.assert_close(got_dc, want["dcost"], rtol=1e-5, atol=1e-7)
        if want["dvar"] is not None:
            torch.testing.assert_close(var.grad, want["dvar"], rtol=1e-5, atol=1e-7)


def test_focal_loss_oracle_on_the_reference_tests_own_cases(golden_dir):
    """tests/modeling/stereo/losses/test_stereo_focal_loss.py of the reference prints its results without asserting
    them; the same two parameterisations (dilated enumeration, per-pixel disp_sample), run through the reference in
    the build container, pin the oracle here."""
    from make_golden import focal_ref_test_inputs
    rec = torch.load(os.path.join(golden_dir, "focal_loss.pt"), weights_only=False)
    for which in (1, 2):
        cost, gt, ds = focal_ref_test_inputs(which)
        cost = cost.clone().requires_grad_(True)
        loss = O.stereo_focal_loss(cost, gt, 2, 5, start_disp=-2, dilation=2 if which == 1 else 1, focal_coefficient=5.0,
                                   disp_sample=ds)
        loss.backward()
        want = rec["ref_test_case%d" % which]
        assert abs(float(loss) - want["loss"]) <= 1e-6 * max(1.0, abs(want["loss"]))
        torch.testing.assert_close(cost.grad, want["dcost"], rtol=1e-5, atol=1e-7)


def test_laplace_disp2prob_oracle_vs_reference(golden_dir):
    """LaplaceDisp2Prob.getProb on the reference's own test cases (tests/modeling/stereo/losses/utils/
    test_disp2prob.py:13-62): dilated enumeration and per-pixel samples; probabilities sum to 1 where the ground truth
    is inside (start, start + max - 1) and are exactly the 1e-40 floor elsewhere."""
    from make_golden import focal_ref_test_inputs
    rec = torch.load(os.path.join(golden_dir, "focal_loss.pt"), weights_only=False)
    for which in (1, 2):
        _, gt, ds = focal_ref_test_inputs(which)
        got = O.laplace_disp2prob(gt.clone(), 5, variance=2, start_disp=-2, dilation=2 if which == 1 else 1, disp_sample=ds)
        want = rec["disp2prob_case%d" % which]
        assert got.shape == want.shape == (1, 3, 3, 4)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=0)
        inside = ((gt > -2) & (gt < 2)).expand_as(got)
        torch.testing.assert_close(got.sum(1)[inside[:, 0]], torch.ones(int(inside[:, 0].sum())), rtol=1e-6, atol=1e-6)
        assert bool((got[~inside] < 1e-39).all())


def test_cmn_eval_matches_reference(golden_dir):
    """The oracle's restatement of the confidence heads against the reference's own Cmn module (cmn.pt)."""
    from make_golden import CMN_CASE, cmn_inputs
    rec = torch.load(os.path.join(golden_dir, "cmn.pt"), weights_only=False)["eval"]
    costs, gt, sd = cmn_inputs()
    confs, cost_vars = O.cmn_eval(sd, costs, CMN_CASE["alpha"], CMN_CASE["beta"])
    for got, want in zip(confs, rec["confs"]):
        torch.testing.assert_close(got, want, atol=1e-6, rtol=1e-5)
    for got, want in zip(cost_vars, rec["cost_vars"]):
        torch.testing.assert_close(got, want, atol=1e-6, rtol=1e-5)

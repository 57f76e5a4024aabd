"""TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/*.pt` by running the REFERENCE ITSELF
(imported read-only from /root/reference through oracle/ref_import.py) on seeded synthetic
inputs.  Run in the build container (the reference tree is absent on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Fixtures are small (inputs are regenerated from seeds at test time; only reference OUTPUTS
and a weight checksum are stored).  The parameterisations include the ones the reference's
own print-style tests use (tests/modeling/stereo/cost_processors/utils/test_cat_fms.py:26-40,
tests/modeling/stereo/disp_predictors/test_disp_predictors.py:42-77).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
import seeded  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

VOLUME_CASES = [
    # name, B, C, H, W, max_disp, start_disp, dilation
    ("ref_test", 1, 1, 3, 4, 5, -2, 2),        # the reference test's own parameterisation
    ("plain", 2, 4, 5, 16, 6, 0, 1),
    ("dilated", 1, 3, 4, 20, 9, 0, 2),
    ("negstart", 1, 2, 3, 12, 7, -3, 1),
    ("odd_lin", 1, 2, 2, 24, 10, 0, 3),        # (max_disp-1) % dilation != 0: truncated linspace
    ("wide", 1, 2, 2, 6, 9, -4, 1),            # |d| reaches W-1 and beyond the image
    ("cfg1ish", 1, 8, 6, 32, 12, 0, 1),
]


def volume_inputs(name, B, C, H, W):
    if name == "ref_test":
        left = torch.linspace(1, H * W, H * W).reshape(1, 1, H, W)
        right = torch.linspace(H * W + 1, H * W * 2, H * W).reshape(1, 1, H, W)
        return left, right
    return seeded.feature_pair(B, C, H, W, seed=len(name) * 7 + C)


def gen_volumes():
    CAT, DIF = ref_import.ref_funcs()
    out = {}
    for name, B, C, H, W, md, sd, dil in VOLUME_CASES:
        l, r = volume_inputs(name, B, C, H, W)
        kw = dict(max_disp=md, start_disp=sd, dilation=dil)
        rec = dict(params=(B, C, H, W, md, sd, dil))
        rec["cat"] = CAT["default"](l, r, **kw)
        rec["dif"] = DIF["default"](l, r, **kw)
        if H > 1 and W > 1 and rec["cat"].shape[2] > 1:
            rec["fast_cat"] = CAT["fast_mode"](l, r, **kw)
            rec["fast_dif"] = DIF["fast_mode"](l, r, **kw)
            rec["fast_dif_norm"] = DIF["fast_mode"](l, r, normalize=True, p=1.0, **kw)
            g = torch.Generator().manual_seed(5)
            D = rec["cat"].shape[2]
            ds = torch.rand(B, D, H, W, generator=g) * md + sd
            rec["disp_sample"] = ds
            rec["fast_cat_sampled"] = CAT["fast_mode"](l, r, disp_sample=ds, **kw)
        out[name] = rec
    torch.save(out, os.path.join(OUT, "volumes.pt"))
    print("volumes.pt", len(out))


PRED_CASES = [
    # name, B, D(max_disp), H, W, start, dilation, alpha, normalize
    ("ref_test_ones", 1, 9, 2, 2, -4, 2, 1.0, True),
    ("plain", 2, 12, 5, 7, 0, 1, 1.0, True),
    ("alpha", 1, 16, 4, 6, 0, 1, 3.0, True),
    ("nonorm", 1, 8, 3, 5, 2, 1, 1.0, False),
    ("dil3", 1, 20, 3, 4, -5, 3, 0.5, True),
    ("d192", 1, 192, 4, 8, 0, 1, 1.0, True),
]


def gen_predictors():
    ref_import.install()
    from dmb.modeling.stereo.disp_predictors.builder import PREDICTORS
    out = {}
    for name, B, md, H, W, sd, dil, alpha, norm in PRED_CASES:
        D = (md + dil - 1) // dil
        if name == "ref_test_ones":
            cost = torch.ones(B, D, H, W)
        else:
            g = torch.Generator().manual_seed(D * 31 + H)
            cost = torch.randn(B, D, H, W, generator=g) * 2.0
        rec = dict(params=(B, md, H, W, sd, dil, alpha, norm), cost=cost)
        kw = dict(max_disp=md, start_disp=sd, dilation=dil, alpha=alpha, normalize=norm)
        rec["DEFAULT"] = PREDICTORS["DEFAULT"](**kw)(cost)
        rec["FASTER"] = PREDICTORS["FASTER"](**kw)(cost).detach()
        g = torch.Generator().manual_seed(3)
        ds = torch.rand(B, D, H, W, generator=g) * md
        rec["disp_sample"] = ds
        rec["DEFAULT_sampled"] = PREDICTORS["DEFAULT"](**kw)(cost, disp_sample=ds)
        for radius, rdil in ((1, 1), (2, 1), (2, 2)):
            rec["LOCAL_r%d_d%d" % (radius, rdil)] = PREDICTORS["LOCAL"](
                radius=radius, radius_dilation=rdil, **kw)(cost)
        out[name] = rec
    torch.save(out, os.path.join(OUT, "predictors.pt"))
    print("predictors.pt", len(out))


def build_ref_processor(agg_type, feat_disp, max_disp):
    cfg = ref_import.load_config("configs/PSMNet/scene_flow.py")
    cfg.model.max_disp = max_disp
    cfg.model.cost_processor.cost_computation.max_disp = feat_disp
    cfg.model.cost_processor.cost_aggregator.max_disp = max_disp
    cfg.model.cost_processor.cost_aggregator.type = agg_type
    cfg.model.disp_predictor.max_disp = max_disp
    proc = ref_import.ref_cost_processor(cfg).eval()
    pred = ref_import.ref_disp_predictor(cfg).eval()
    return proc, pred


def gen_aggregators():
    """Config 1: features [1,32,16,32], cost_computation.max_disp=12,
    aggregator/predictor max_disp=48.  Seeded weights (plain and sharpened), randomised BN."""
    out = {}
    for agg_type in ("PSMNet", "AcfNet"):
        proc, pred = build_ref_processor(agg_type, 12, 48)
        entries = seeded.aggregator_entries(agg_type, 64)
        ref_sd = proc.aggregator.state_dict()
        assert [k for k, _, _ in entries] == list(ref_sd.keys()), "state-dict key order differs from reference"
        for k, shape, _ in entries:
            assert tuple(ref_sd[k].shape) == tuple(shape), (k, ref_sd[k].shape, shape)
        for variant, seed, sharpen, shift in (("plain", 0, 1.0, None), ("sharp", 1, 4.0, 5)):
            sd = seeded.seeded_state_dict(entries, seed=seed, sharpen=sharpen)
            proc.aggregator.load_state_dict(sd)
            l, r = seeded.feature_pair(1, 32, 16, 32, seed=100 + seed, scale=0.5, shift=shift)
            with torch.no_grad():
                costs = proc(l, r)
                disps = [pred(c) for c in costs]
            rec = dict(seed=seed, sharpen=sharpen, shift=shift, weight_checksum=seeded.checksum(sd),
                       disps=[d.clone() for d in disps],
                       # strided sample of the full-resolution costs (keeps the fixture small)
                       cost_samples=[c[:, ::3, ::4, ::4].clone() for c in costs],
                       cost_absmax=[float(c.abs().max()) for c in costs])
            out["%s_%s" % (agg_type, variant)] = rec
            print(agg_type, variant, "disp mean/std", float(disps[0].mean()), float(disps[0].std()))
    torch.save(out, os.path.join(OUT, "aggregators.pt"))
    print("aggregators.pt", len(out))


TRAIN_CASE = dict(B=2, C=32, H4=8, W4=16, feat_disp=8, max_disp=32, seed=3, feat_seed=21, shift=2)


def train_inputs():
    """Seeded inputs of the training-step fixture (regenerated identically at test time)."""
    c = TRAIN_CASE
    l, r = seeded.feature_pair(c["B"], c["C"], c["H4"], c["W4"], seed=c["feat_seed"], scale=0.5, shift=c["shift"])
    g = torch.Generator().manual_seed(c["feat_seed"] + 1)
    gt = torch.rand(c["B"], 1, c["H4"] * 4, c["W4"] * 4, generator=g) * (c["max_disp"] + 8) - 4   # some masked out
    return l, r, gt


def grad_summary(t, n=32):
    """(sum, abs-sum, strided sample) of a gradient tensor: keeps the fixture small."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return dict(sum=float(f.double().sum()), abssum=float(f.double().abs().sum()), sample=f[::step][:n].clone(),
                step=step)


def gen_train_step():
    """One training step (forward with batch-statistics BatchNorm, smooth-L1 loss, backward) of the
    REFERENCE cost processor + predictor: loss, disparities, gradient summaries, running statistics."""
    ref_import.install()
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_smooth_l1", os.path.join(ref_import.REFERENCE_ROOT, "dmb/modeling/stereo/losses/smooth_l1_loss.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    c = TRAIN_CASE
    out = {}
    for agg_type in ("PSMNet", "AcfNet"):
        proc, pred = build_ref_processor(agg_type, c["feat_disp"], c["max_disp"])
        proc.train(); pred.train()
        entries = seeded.aggregator_entries(agg_type, 64)
        sd = seeded.seeded_state_dict(entries, seed=c["seed"])
        proc.aggregator.load_state_dict(sd)
        l, r, gt = train_inputs()
        l.requires_grad_(True); r.requires_grad_(True)
        costs = proc(l, r)
        disps = [pred(cost) for cost in costs]
        losses = mod.DispSmoothL1Loss(c["max_disp"], weights=(1.0, 0.7, 0.5))(disps, gt)
        loss = sum(losses.values())
        loss.backward()
        grads = {k: grad_summary(p.grad if p.grad is not None else torch.zeros_like(p))
                 for k, p in proc.aggregator.named_parameters()}
        after = proc.aggregator.state_dict()
        running = {k: v.clone() for k, v in after.items() if "running_" in k or "num_batches" in k}
        out[agg_type] = dict(case=dict(c), weight_checksum=seeded.checksum(sd), loss=float(loss),
                             disps=[d.detach().clone() for d in disps], grads=grads,
                             dleft=l.grad.clone(), dright=r.grad.clone(), running=running)
        print("train", agg_type, "loss", float(loss), "|dleft|", float(l.grad.abs().sum()))
    torch.save(out, os.path.join(OUT, "train_step.pt"))
    print("train_step.pt")


def gen_hourglass():
    """A bare Hourglass (cost_processors/utils/hourglass.py) with presqu/postsqu given."""
    ref_import.install()
    from dmb.modeling.stereo.cost_processors.utils.hourglass import Hourglass
    hg = Hourglass(in_planes=32, batch_norm=True).eval()
    entries = []
    seeded._hourglass(entries, "hg", 32, bias=False)
    entries = [(k[3:], s, r) for k, s, r in entries]
    sd = seeded.seeded_state_dict(entries, seed=7)
    assert list(sd.keys()) == list(hg.state_dict().keys())
    hg.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 32, 8, 8, 12, generator=g)
    pre = torch.randn(1, 64, 4, 4, 6, generator=g)
    post = torch.randn(1, 64, 4, 4, 6, generator=g)
    with torch.no_grad():
        o1 = hg(x, None, None)
        o2 = hg(x, pre, post)
    torch.save(dict(weight_checksum=seeded.checksum(sd), first=[t.clone() for t in o1],
                    second=[t.clone() for t in o2]), os.path.join(OUT, "hourglass.pt"))
    print("hourglass.pt")


FOCAL_CASES = [
    # name, B, D (= samples), cost H, W, gt H, W, max_disp, start_disp, dilation, coefficient, sparse, variance kind
    ("full_res_fc0_scalar_var", 2, 24, 8, 16, 8, 16, 24, 0, 1, 0.0, False, "scalar"),
    ("full_res_fc5_tensor_var", 2, 24, 8, 16, 8, 16, 24, 0, 1, 5.0, False, "tensor"),      # AcfNet adaptive
    ("half_res_fc5_tensor_var", 2, 12, 8, 16, 16, 32, 24, 0, 1, 5.0, False, "tensor"),     # cost at 1/2 of the gt
    ("sparse_fc2", 1, 16, 6, 10, 12, 20, 32, 0, 1, 2.0, True, "scalar"),                   # KITTI-style zeros
    ("start_dilation", 1, 10, 5, 7, 5, 7, 20, -4, 2, 1.0, False, "tensor"),
    ("all_invalid", 1, 8, 4, 5, 4, 5, 8, 0, 1, 5.0, False, "scalar"),
]


def focal_inputs(case):
    name, B, D, H, W, Hg, Wg, max_disp, start, dil, fc, sparse, vkind = case
    g = torch.Generator().manual_seed(len(name) * 7 + D)
    cost = torch.randn(B, D, H, W, generator=g) * 3.0
    gt = torch.rand(B, 1, Hg, Wg, generator=g) * (max_disp + 6) + start - 3        # some outside the range
    if sparse:
        gt = gt * (torch.rand(B, 1, Hg, Wg, generator=g) > 0.5)
    if name == "all_invalid":
        gt = torch.full((B, 1, Hg, Wg), 1000.0)
    var = 1.2 if vkind == "scalar" else torch.rand(B, 1, H, W, generator=g) * 1.5 + 0.5
    return cost, gt, var


def focal_ref_test_inputs(which):
    """The reference's own test parameterisations (tests/modeling/stereo/losses/test_stereo_focal_loss.py:17-63 and
    :65-110): max_disp 5, start -2, h x w = 3 x 4, variance 2, coefficient 5, an all-ones cost volume; case 1 enumerates
    the samples with dilation 2, case 2 hands them in per pixel as disp_sample = [-2, 0, 2].  (The reference draws its
    ground truth with the global RNG; seeded here.)"""
    g = torch.Generator().manual_seed(40 + which)
    gt = torch.rand(1, 1, 3, 4, generator=g) * 5 - 2
    cost = torch.ones(1, 3, 3, 4)
    ds = None
    if which == 2:
        ds = torch.tensor([-2.0, 0.0, 2.0]).repeat(1, 3, 4, 1).permute(0, 3, 1, 2).contiguous()
    return cost, gt, ds


def gen_focal_loss():
    """StereoFocalLoss of the reference (losses/stereo_focal_loss.py) on seeded inputs: loss, d/dcost, d/dvariance."""
    ref_import.install()
    from dmb.modeling.stereo.losses.stereo_focal_loss import StereoFocalLoss
    out = {}
    for case in FOCAL_CASES:
        name, B, D, H, W, Hg, Wg, max_disp, start, dil, fc, sparse, vkind = case
        cost, gt, var = focal_inputs(case)
        cost = cost.clone().requires_grad_(True)
        if torch.is_tensor(var):
            var = var.clone().requires_grad_(True)
        ev = StereoFocalLoss(max_disp=max_disp, start_disp=start, dilation=dil, weights=(0.7,), focal_coefficient=fc,
                             sparse=sparse)
        loss = ev(cost, gt, var)["stereo_focal_loss_lvl0"]
        if loss.requires_grad:
            loss.backward()
        out[name] = dict(loss=float(loss), dcost=cost.grad.clone() if cost.grad is not None else torch.zeros_like(cost),
                         dvar=(var.grad.clone() if torch.is_tensor(var) and var.grad is not None else None))
        print("focal", name, float(loss))
    for which in (1, 2):
        cost, gt, ds = focal_ref_test_inputs(which)
        cost = cost.clone().requires_grad_(True)
        ev = StereoFocalLoss(max_disp=5, start_disp=-2, dilation=2 if which == 1 else 1, weights=(1.0), focal_coefficient=5.0,
                             sparse=False)
        loss = ev(estCost=cost, gtDisp=gt, variance=2, disp_sample=ds)["stereo_focal_loss_lvl0"]
        loss.backward()
        out["ref_test_case%d" % which] = dict(loss=float(loss), dcost=cost.grad.clone(), dvar=None)
        print("focal ref_test_case%d" % which, float(loss))
    # the probability volumes of tests/modeling/stereo/losses/utils/test_disp2prob.py:13-62 (same two cases)
    from dmb.modeling.stereo.losses.utils.disp2prob import LaplaceDisp2Prob
    for which in (1, 2):
        _, gt, ds = focal_ref_test_inputs(which)
        prob = LaplaceDisp2Prob(gt.clone(), max_disp=5, variance=2, start_disp=-2, dilation=2 if which == 1 else 1,
                                disp_sample=ds).getProb()
        out["disp2prob_case%d" % which] = prob.clone()
    torch.save(out, os.path.join(OUT, "focal_loss.pt"))
    print("focal_loss.pt")


OTHER_AGG_CASES = dict(
    GCNet=dict(in_planes=64, dims=(16, 16, 32), seed=23, max_disp=32),
    StereoNet=dict(in_planes=32, dims=(6, 10, 34), seed=29, max_disp=6),
)


def other_agg_input(kind):
    c = OTHER_AGG_CASES[kind]
    g = torch.Generator().manual_seed(c["seed"] + 100)
    return torch.randn(1, c["in_planes"], *c["dims"], generator=g) * 0.5


def gen_other_aggregators():
    """GCAggregator / StereoNetAggregator of the reference itself (aggregators/GCNet.py:7-120, StereoNet.py:9-55) in eval
    mode with seeded weights (seeded.module_entries on the REFERENCE module): output volumes + an entry checksum."""
    from dmb.modeling.stereo.cost_processors.aggregators.GCNet import GCAggregator
    from dmb.modeling.stereo.cost_processors.aggregators.StereoNet import StereoNetAggregator
    out = {}
    for kind, cls in (("GCNet", GCAggregator), ("StereoNet", StereoNetAggregator)):
        c = OTHER_AGG_CASES[kind]
        m = cls(max_disp=c["max_disp"], in_planes=c["in_planes"], batch_norm=True)
        entries = seeded.module_entries(m)
        sd = seeded.seeded_state_dict(entries, seed=c["seed"])
        m.load_state_dict(sd)
        m.eval()
        with torch.no_grad():
            y = m(other_agg_input(kind))[0]
        out[kind] = dict(cost=y.clone(), weight_checksum=seeded.checksum(sd), n_entries=len(entries))
        print("other aggregator", kind, tuple(y.shape), float(y.abs().max()))
    torch.save(out, os.path.join(OUT, "other_aggregators.pt"))


CMN_CASE = dict(in_planes=192, num=2, alpha=1.0, beta=1.0, B=2, H=10, W=36, seed=17)


def cmn_inputs():
    """Seeded inputs of the Cmn fixture: `num` cost volumes [B,192,H,W] with a wide value range, a ground-truth map with
    some invalid pixels, and a state dict for reference-layout keys (weights N(0, 0.05), BatchNorm statistics)."""
    c = CMN_CASE
    g = torch.Generator().manual_seed(c["seed"])
    costs = [torch.randn(c["B"], c["in_planes"], c["H"], c["W"], generator=g) * 3.0 for _ in range(c["num"])]
    gt = torch.rand(c["B"], 1, c["H"], c["W"], generator=g) * 230 - 10
    sd = {}
    mid = c["in_planes"] // 3
    for i in range(c["num"]):
        p = "conf_heads.%d.conf_net." % i
        sd[p + "0.0.weight"] = torch.randn(mid, c["in_planes"], 3, 3, generator=g) * 0.05
        sd[p + "0.1.weight"] = torch.rand(mid, generator=g) + 0.5
        sd[p + "0.1.bias"] = torch.randn(mid, generator=g) * 0.2
        sd[p + "0.1.running_mean"] = torch.randn(mid, generator=g) * 0.3
        sd[p + "0.1.running_var"] = torch.rand(mid, generator=g) + 0.5
        sd[p + "0.1.num_batches_tracked"] = torch.tensor(3)
        sd[p + "1.weight"] = torch.randn(1, mid, 1, 1, generator=g) * 0.2
    return costs, gt, sd


def gen_cmn():
    """The reference's own Cmn (cmn/cmn.py:40-82 + cmn/loss.py + losses/conf_nll_loss.py) in eval and train mode."""
    from dmb.modeling.stereo.cmn.cmn import Cmn
    c = CMN_CASE
    cfg = ref_import.ConfigDict(model=dict(batch_norm=True, cmn=dict(
        in_planes=c["in_planes"], num=c["num"], alpha=c["alpha"], beta=c["beta"],
        losses=dict(nll_loss=dict(max_disp=192, weights=(1.0, 0.7), weight=8.0)))), data=dict(sparse=False))
    costs, gt, sd = cmn_inputs()
    out = {}
    m = Cmn(cfg, c["in_planes"], c["num"], c["alpha"], c["beta"])
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        cost_vars, confs = m([t.clone() for t in costs], gt)
    out["eval"] = dict(cost_vars=[t.clone() for t in cost_vars], confs=[t.clone() for t in confs])
    m.train()
    xs = [t.clone().requires_grad_(True) for t in costs]
    cost_vars, losses = m(xs, gt)
    total = sum(losses.values()) + sum(v.mean() for v in cost_vars)
    total.backward()
    out["train"] = dict(losses={k: float(v) for k, v in losses.items()}, cost_vars=[t.detach().clone() for t in cost_vars],
                        dcost=[grad_summary(x.grad) for x in xs],
                        grads={k: grad_summary(p.grad) for k, p in m.named_parameters()},
                        running={k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k})
    torch.save(out, os.path.join(OUT, "cmn.pt"))
    print("cmn.pt", out["train"]["losses"])


def gen_epe():
    # the package __init__ chain pulls visualisation deps; load the single file instead
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_pixel_error", os.path.join(ref_import.REFERENCE_ROOT, "dmb/data/datasets/evaluation/stereo/pixel_error.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    calc_error = mod.calc_error
    g = torch.Generator().manual_seed(2)
    gt = torch.rand(1, 1, 16, 24, generator=g) * 220 - 10
    est = gt + torch.randn(1, 1, 16, 24, generator=g)
    e = calc_error(est, gt, 0, 192)
    torch.save(dict(gt=gt, est=est, epe=float(e["epe"])), os.path.join(OUT, "epe.pt"))
    print("epe.pt", float(e["epe"]))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    gen_volumes()
    gen_predictors()
    gen_hourglass()
    gen_epe()
    gen_aggregators()
    gen_train_step()
    gen_focal_loss()
    gen_cmn()
    gen_other_aggregators()

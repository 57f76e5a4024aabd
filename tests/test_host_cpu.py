"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, argument
validation returns error codes (no GPU needed), and the host-side logic (registry tables, config
handling, state-dict layout, BN folding, weight packing, DeferredCost dispatch)."""
import os
import re

import pytest
import torch
import torch.nn.functional as F

import dmb_oracle as O
import seeded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import __graft_entry__
    __graft_entry__.build()
    import densematchingbenchmark_b200 as pkg
    return pkg


def test_library_exports_every_declared_symbol(P):
    from densematchingbenchmark_b200 import _cabi as C
    lib = C.load()
    header = open(os.path.join(ROOT, "include", "dmb_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dmb_b200_\w+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    bound = set(C.SIGNATURES) | set(C.OTHER)
    assert set(declared) == bound, "ctypes table and header disagree: %s" % (set(declared) ^ bound)
    assert lib.dmb_b200_abi_version() >= 1


def test_argument_validation_without_gpu(P):
    from densematchingbenchmark_b200 import _cabi as C
    with pytest.raises(C.DmbB200Error) as e:
        C.call("dmb_b200_cat_volume", None, None, None, 1, 4, 4, 4, C.int_array([0]), 1, None)
    assert "null pointer" in str(e.value)
    with pytest.raises(C.DmbB200Error):
        C.call("dmb_b200_upsample_regress", None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, 1.0, 1, 0.0, 1.0, None, None)
    with pytest.raises(C.DmbB200Error):   # CPU tensors are rejected by the binding: no CPU path
        P.CAT_FUNCS["default"](torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 3, 4), max_disp=2)


def test_tables_cover_reference_keys(P):
    assert {"Concatenation", "Difference", "Correlation", "GroupWiseCorrelation"} <= set(P.PROCESSORS)
    assert set(P.CAT_FUNCS) == {"default", "fast_mode"} and set(P.DIF_FUNCS) == {"default", "fast_mode"}
    assert {"PSMNet", "AcfNet", "GCNet", "StereoNet"} <= set(P.AGGREGATORS)
    assert set(P.PREDICTORS) == {"DEFAULT", "FASTER", "LOCAL"}


def _cfg(P, agg="PSMNet"):
    return P.ConfigDict(model=dict(
        batch_norm=True,
        cost_processor=dict(type="Concatenation",
                            cost_computation=dict(type="default", max_disp=12, start_disp=0, dilation=1),
                            cost_aggregator=dict(type=agg, max_disp=48, in_planes=64)),
        disp_predictor=dict(type="FASTER", max_disp=48, start_disp=0, dilation=1, alpha=1.0, normalize=True)))


@pytest.mark.parametrize("agg", ["PSMNet", "AcfNet"])
def test_state_dict_layout_matches_reference(P, agg):
    proc = P.build_cost_processor(_cfg(P, agg))
    entries = seeded.aggregator_entries(agg, 64)
    sd = proc.aggregator.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in entries]
    for k, shape, _ in entries:
        assert tuple(sd[k].shape) == tuple(shape)
    proc.aggregator.load_state_dict(seeded.seeded_state_dict(entries, 0))     # loads unchanged
    pred = P.build_disp_predictor(_cfg(P, agg))
    assert list(pred.state_dict().keys()) == ["disp_regression.weight"]
    assert torch.equal(pred.state_dict()["disp_regression.weight"].reshape(-1), O.disp_samples(48))
    assert not pred.disp_regression.weight.requires_grad


def test_builder_errors_like_reference(P):
    cfg = _cfg(P)
    cfg.model.cost_processor.type = "Nope"
    with pytest.raises(AssertionError):
        P.build_cost_processor(cfg)
    cfg = _cfg(P)
    cfg.model.cost_processor.cost_aggregator.type = "Nope"
    with pytest.raises(AssertionError):
        P.build_cost_processor(cfg)
    cfg = _cfg(P)
    cfg.model.disp_predictor.type = "Nope"
    with pytest.raises(AssertionError):
        P.build_disp_predictor(cfg)
    # building must not consume the caller's config (reference copies before pop, builder.py:26-27)
    cfg = _cfg(P)
    P.build_cost_processor(cfg)
    assert cfg.model.cost_processor.cost_computation.type == "default"


def test_disp_indices_match_oracle(P):
    from densematchingbenchmark_b200.ops import functional as F_
    for md, sd, dil in ((48, 0, 1), (5, -2, 2), (10, 0, 3), (9, -4, 1), (192, 0, 1), (7, 3, 4)):
        assert F_.disp_indices(md, sd, dil) == O.disp_indices(md, sd, dil)


def test_bn_folding_and_weight_packing(P):
    from densematchingbenchmark_b200.modeling.stereo.layers.basic_layers import conv3d_bn_relu, deconv3d_bn
    torch.manual_seed(0)
    for unit, transposed in ((conv3d_bn_relu(True, 6, 5, 3, 1, 1, bias=True), False),
                             (deconv3d_bn(True, 6, 4, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False), True)):
        unit.eval()
        unit[1].running_mean.normal_(0, 0.2); unit[1].running_var.uniform_(0.5, 2.0)
        unit[1].weight.data.uniform_(0.5, 1.5); unit[1].bias.data.normal_(0, 0.1)
        x = torch.randn(1, 6, 4, 5, 6)
        with torch.no_grad():
            conv = unit[0]
            if transposed:
                y = F.conv_transpose3d(x, conv.weight, conv.bias, stride=2, padding=1, output_padding=1)
            else:
                y = F.conv3d(x, conv.weight, conv.bias, padding=1)
            want = unit[1](y)
        wp, b = unit.folded()
        K3, Cin, Cout = wp.shape
        # un-pack and apply with torch to check the folded parameters themselves
        if transposed:
            w = wp.reshape(3, 3, 3, Cin, Cout).permute(3, 4, 0, 1, 2)
            got = F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1)
        else:
            w = wp.reshape(3, 3, 3, Cin, Cout).permute(4, 3, 0, 1, 2)
            got = F.conv3d(x, w, b, padding=1)
        torch.testing.assert_close(got, want, atol=1e-5, rtol=1e-5)
        again = unit.folded()
        assert again[0] is wp                       # cached
        unit[1].running_mean.add_(1.0)              # in-place change invalidates the cache
        assert unit.folded()[0] is not wp


def test_deferred_cost_dispatch(P, monkeypatch):
    """Any op other than our predictors materialises the dense cost exactly once."""
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import deferred
    calls = {"n": 0}

    def fake_upsample(low, out_dhw, mode="trilinear", up_weight=None, want_cost=True, want_disp=False, **kw):
        calls["n"] += 1
        cost = F.interpolate(low.unsqueeze(1), list(out_dhw), mode="trilinear", align_corners=True).squeeze(1)
        return (cost if want_cost else None), (O.soft_argmin(cost, out_dhw[0]) if want_disp else None)

    monkeypatch.setattr(deferred.F_, "upsample_regress", fake_upsample)
    low = torch.randn(1, 3, 4, 5)
    dc = deferred.DeferredCost(low, (12, 16, 20), "trilinear")
    assert dc.dim() == 4 and tuple(dc.shape) == (1, 12, 16, 20) and dc.shape[1] == 12
    want = F.interpolate(low.unsqueeze(1), [12, 16, 20], mode="trilinear", align_corners=True).squeeze(1)
    torch.testing.assert_close(dc.regress(), O.soft_argmin(want, 12))        # fused route, nothing materialised
    assert dc._dense is None
    torch.testing.assert_close(torch.softmax(dc, dim=1), torch.softmax(want, dim=1))
    torch.testing.assert_close(dc[:, ::2] * 2.0, want[:, ::2] * 2.0)
    torch.testing.assert_close(dc.cpu().clone(), want)
    assert calls["n"] == 2       # one regress + one materialisation


def test_training_forward_has_no_cpu_path_either(P):
    """train() routes through the autograd Functions of ops/autograd.py -- CUDA kernels only."""
    from densematchingbenchmark_b200._cabi import DmbB200Error
    from densematchingbenchmark_b200.modeling.stereo.layers.basic_layers import conv3d_bn
    unit = conv3d_bn(True, 4, 4, 3, 1, 1).train()
    with pytest.raises(DmbB200Error):
        unit(torch.zeros(1, 4, 2, 2, 2))


def test_spn_module_rejects_cpu(P):
    from densematchingbenchmark_b200.ops import GateRecurrent2dnoind
    z = torch.zeros(1, 1, 2, 2)
    with pytest.raises(RuntimeError):
        GateRecurrent2dnoind(True, False)(z, z, z, z)


@pytest.mark.ref
def test_dropin_swaps_reference_tables(P):
    """With the reference importable (build container only), install_into_dmb() makes the stock
    config resolve to our classes and the reference checkpoint layout still loads."""
    import ref_import
    ref_import.install()
    replaced = P.install_into_dmb("dmb")
    assert "PSMNet" in replaced["AGGREGATORS"]
    cfg = ref_import.load_config("configs/PSMNet/scene_flow.py")
    from dmb.modeling.stereo.cost_processors import build_cost_processor
    from dmb.modeling.stereo.disp_predictors import build_disp_predictor
    proc = build_cost_processor(cfg)
    assert type(proc).__module__.startswith("densematchingbenchmark_b200")
    assert type(proc.aggregator).__module__.startswith("densematchingbenchmark_b200")
    assert type(build_disp_predictor(cfg)).__module__.startswith("densematchingbenchmark_b200")
    # the loss builder of the AcfNet config now constructs (and dispatches on) our fused StereoFocalLoss
    assert replaced.get("losses") == ["StereoFocalLoss"]
    acf = ref_import.load_config("configs/AcfNet/scene_flow_adaptive.py")
    from dmb.modeling.stereo.losses.builder import make_focal_loss_evaluator
    ev = make_focal_loss_evaluator(acf)
    assert type(ev).__module__.startswith("densematchingbenchmark_b200")
    assert (ev.max_disp, ev.focal_coefficient, tuple(ev.weights)) == (192, 5.0, (1.0, 0.7, 0.5))


def test_tc_schedule_tiles_and_depth_segments(P):
    """Host arithmetic of the tcgen05 launches (no GPU): the 4x32 stride-1 tile (30 output columns per step) tiles
    the PSMNet grids exactly, depth segments cover every output plane exactly once, and the slowest persistent CTA
    of the static schedule stays close to the average load at the config-2 sizes."""
    import ctypes
    from densematchingbenchmark_b200 import _cabi as C

    def sched(kind, B, D, H, W):
        out = (ctypes.c_int * 8)()
        C.call("dmb_b200_conv3d_tc_schedule", kind, B, D, H, W, out)
        return dict(zip(("tiles_h", "tiles_w", "nseg", "seg_len", "items", "grid", "th", "tw"), list(out)))

    # exact tiling of the 1/4, 1/8, 1/16 grids of a 544x960 image by the stride-1 kernel (kind 3)
    for (D, H, W), tiles in (((48, 136, 240), (34, 8)), ((24, 68, 120), (17, 4)), ((12, 34, 60), (9, 2))):
        s3 = sched(3, 1, D, H, W)
        assert (s3["th"], s3["tw"]) == (4, 30)
        assert (s3["tiles_h"], s3["tiles_w"]) == tiles
        assert s3["tiles_w"] * 30 == W                                  # no ragged last tile column
    for kind in (0, 1, 2, 3, 4, 5):
        for B in (1, 2, 3):
            for (D, H, W) in ((48, 136, 240), (24, 68, 120), (12, 34, 60), (6, 10, 14), (2, 4, 8), (16, 32, 64)):
                s_ = sched(kind, B, D, H, W)
                Dm = D // 2 if kind in (1, 4) else D
                assert s_["nseg"] >= 1 and s_["seg_len"] >= 1
                assert s_["nseg"] * s_["seg_len"] >= Dm > (s_["nseg"] - 1) * s_["seg_len"]      # a partition of the planes
                nb = 1 if kind == 1 else B
                assert s_["items"] == s_["tiles_h"] * s_["tiles_w"] * nb * s_["nseg"]
                assert 1 <= s_["grid"] <= 148 and s_["grid"] == min(s_["items"], 148)
    # load balance of the static schedule where it matters (the 1/4-resolution layers, batch 1 and 2)
    for B in (1, 2):
        s3 = sched(3, B, 48, 136, 240)
        slowest = -(-s3["items"] // s3["grid"]) * s3["seg_len"]
        average = s3["tiles_h"] * s3["tiles_w"] * B * 48 / 148.0
        assert slowest <= 1.10 * average, (slowest, average)


def test_cached_weight_scale_refreshes_with_parameter_updates(P):
    """The fp16 weight pre-scale of the training path is a power of two, cached on the Parameter and recomputed only
    after `refresh` in-place updates (no per-step host synchronisation)."""
    import torch
    from densematchingbenchmark_b200.modeling.stereo.cost_processors.aggregators import tc_engine as T
    w = torch.nn.Parameter(torch.full((4, 4), 0.3))
    s0 = T.cached_weight_scale(w, w.detach(), refresh=3)
    assert s0 == 16.0                                                   # 0.3 * 16 = 4.8 in [4, 8)
    with torch.no_grad():
        w.mul_(100.0)                                                   # one update: still the cached value
    assert T.cached_weight_scale(w, w.detach(), refresh=3) == s0
    with torch.no_grad():
        w.add_(0.0); w.add_(0.0)                                        # third update since the scale was taken
    assert T.cached_weight_scale(w, w.detach(), refresh=3) == 0.25      # 30 * 0.25 = 7.5 in [4, 8)


@pytest.mark.ref
def test_reference_spn_kernel_compiles_into_oracle_ref():
    """oracle/build_ref.py compiles the reference's own SPN CUDA file from where it lies (build container only);
    the resulting test-only library exports the two C entry points tests/test_gpu_spn_ref.py binds."""
    import ctypes
    import build_ref
    lib = build_ref.build_ref()
    assert lib is not None and os.path.exists(lib)
    assert os.path.realpath(lib).startswith(os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref")))
    h = ctypes.CDLL(lib)
    assert hasattr(h, "spn_ref_forward") and hasattr(h, "spn_ref_backward")


def test_backbone_train_mode_runs_the_two_views_separately():
    """Reference backbones/PSMNet.py:126-127 runs `_forward` once per view: in train() the BatchNorm statistics are
    per view and the running statistics get two updates.  Eval mode may batch the views (same result)."""
    from densematchingbenchmark_b200.modeling.stereo.backbones.PSMNet import PSMNetBackbone
    torch.manual_seed(0)
    bb = PSMNetBackbone(3, True)
    l, r = torch.randn(2, 3, 256, 256), torch.randn(2, 3, 256, 256)     # branch1 pools 64x64 windows of the 1/4 map
    import copy
    twin = copy.deepcopy(bb)
    bb.train(); twin.train()
    fl, fr = bb(l, r)
    wl, wr = twin._forward(l), twin._forward(r)
    assert torch.equal(fl, wl) and torch.equal(fr, wr)
    bns = [m for m in bb.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert all(int(m.num_batches_tracked) == 2 for m in bns)
    bb.eval(); twin.eval()
    with torch.no_grad():
        el, er = bb(l, r)
        torch.testing.assert_close(el, twin._forward(l), atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(er, twin._forward(r), atol=1e-5, rtol=1e-5)


def test_cmn_mirror_contract(P):
    """Cmn / ConfHead mirror (cmn/cmn.py:10-93): reference constructor, state-dict keys, and no CPU fallback."""
    from densematchingbenchmark_b200.modeling.stereo.cmn import build_cmn
    from densematchingbenchmark_b200 import _cabi
    cfg = P.ConfigDict(model=dict(batch_norm=True, cmn=dict(in_planes=192, num=3, alpha=1.0, beta=1.0,
                                                            losses=dict(nll_loss=dict(max_disp=192, weights=(1.0, 0.7, 0.5), weight=8.0)))),
                       data=dict(sparse=False))
    m = build_cmn(cfg)
    keys = set(m.state_dict().keys())
    for i in range(3):
        for suffix in ("0.0.weight", "0.1.weight", "0.1.bias", "0.1.running_mean", "0.1.running_var", "0.1.num_batches_tracked",
                       "1.weight"):
            assert "conf_heads.%d.conf_net.%s" % (i, suffix) in keys
    assert tuple(m.conf_heads[0].conf_net[0][0].weight.shape) == (64, 192, 3, 3)
    with pytest.raises(_cabi.DmbB200Error):
        m.eval()([torch.zeros(1, 192, 4, 8)] * 3)
    with pytest.raises(AssertionError):
        m.get_confidence([torch.zeros(1, 192, 4, 8)])

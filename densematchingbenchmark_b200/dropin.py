"""Drop the B200 implementations into an importable reference `dmb` package.

`dmb` resolves components through plain module-level dicts: replacing the
dict VALUES (and adding the new keys) is all it takes for `GeneralizedStereoModel` built from the
stock PSMNet / AcfNet / GCNet / StereoNet configs to run this path, with reference checkpoints
loading unchanged (identical state-dict keys).  See INTEGRATION.md."""
import importlib


def install_into_dmb(dmb_module_name="dmb"):
    from .modeling.stereo import cost_processors as cp
    from .modeling.stereo import disp_predictors as dp
    ref_cp = importlib.import_module(dmb_module_name + ".modeling.stereo.cost_processors.builder")
    ref_agg = importlib.import_module(dmb_module_name + ".modeling.stereo.cost_processors.aggregators.builder")
    ref_cat = importlib.import_module(dmb_module_name + ".modeling.stereo.cost_processors.utils.cat_fms")
    ref_dif = importlib.import_module(dmb_module_name + ".modeling.stereo.cost_processors.utils.dif_fms")
    ref_dp = importlib.import_module(dmb_module_name + ".modeling.stereo.disp_predictors.builder")
    replaced = {}
    for ours, theirs, name in ((cp.CAT_FUNCS, ref_cat.CAT_FUNCS, "CAT_FUNCS"),
                               (cp.DIF_FUNCS, ref_dif.DIF_FUNCS, "DIF_FUNCS"),
                               (cp.AGGREGATORS, ref_agg.AGGREGATORS, "AGGREGATORS"),
                               (dp.PREDICTORS, ref_dp.PREDICTORS, "PREDICTORS")):
        theirs.update(ours)
        replaced[name] = sorted(ours.keys())
    # processors: ours resolve the volume function at construction through OUR tables and build
    # the aggregator through OUR builder, so swap the classes (keeps 'DeepPruner'/'AnyNet' as is)
    # ('Correlation': the reference's own COR_FUNCS import needs the un-vendored spatial_correlation_sampler; ours does not)
    for key in ("Concatenation", "Difference", "Correlation", "GroupWiseCorrelation"):
        ref_cp.PROCESSORS[key] = cp.PROCESSORS[key]
    replaced["PROCESSORS"] = ["Concatenation", "Difference", "Correlation", "GroupWiseCorrelation"]
    # dmb.ops.spn
    try:
        ref_ops = importlib.import_module(dmb_module_name + ".ops")
        from .ops.spn import GateRecurrent2dnoind
        ref_ops.GateRecurrent2dnoind = GateRecurrent2dnoind
        replaced["ops"] = ["GateRecurrent2dnoind"]
    except Exception:   # the reference's dmb.ops import needs its compiled extension
        pass
    # losses: the builder constructs `StereoFocalLoss` and dispatches on isinstance(..., StereoFocalLoss) through its
    # module-level name (dmb/modeling/stereo/losses/builder.py:3,39,70) -- rebinding that name swaps both
    try:
        ref_loss = importlib.import_module(dmb_module_name + ".modeling.stereo.losses.builder")
        from .modeling.stereo.losses import StereoFocalLoss
        ref_loss.StereoFocalLoss = StereoFocalLoss
        replaced["losses"] = ["StereoFocalLoss"]
    except Exception:
        pass
    # confidence measurement network: GeneralizedStereoModel.__init__ calls the module-level name `build_cmn`
    # (dmb/modeling/stereo/models/general_stereo_model.py:8,36-38)
    try:
        ref_gsm = importlib.import_module(dmb_module_name + ".modeling.stereo.models.general_stereo_model")
        from .modeling.stereo.cmn import build_cmn
        ref_gsm.build_cmn = build_cmn
        replaced["cmn"] = ["build_cmn"]
    except Exception:
        pass
    return replaced

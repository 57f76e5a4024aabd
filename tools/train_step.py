"""BASELINE config 5: one data-parallel training step of the stereo model (AcfNet / PSMNet aggregator) on
synthetic SceneFlow-sized crops -- backbone (plain torch autograd, outside the hot path) + cat volume +
aggregator (training-mode BatchNorm, synchronised over the ranks) + soft-argmin, the configuration's losses
(AcfNet: stereo focal loss on the three cost volumes + 0.1 x smooth-L1 on the disparities,
configs/AcfNet/scene_flow_uniform.py:52-79; PSMNet: smooth-L1, configs/PSMNet/scene_flow.py:55-63), backward through
the library's own kernels, bucketed NCCL gradient all-reduce overlapped with backward (the ONE collective on the
path: dmb/utils/dist_utils.py:16-47), gradient clipping, RMSprop step.

    python tools/train_step.py                                   # 1 GPU, 4 pairs of 256x512, D=192
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_step.py --sync-bn                            # 8 x 4 pairs

Prints one JSON line (rank 0): ms per step (max over ranks, CUDA events), pairs/s over all ranks, and the
split forward / backward(+overlapped all-reduce) / reducer tail / optimizer.  `run_train_bench` is the same
measurement as a function: bench.py calls it for the `train` block of its JSON line.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_train_bench(kind="AcfNet", batch=4, height=256, width=512, max_disp=192, steps=3, warmup=2, sync_bn=True,
                    backbone=True, bucket_mb=4.0, loss="config", rank=0, world=1, device=None, sync_backbone_bn=True,
                    channels_last_backbone=True, graph_backbone=True, graph_hot_path=False):
    """Times `steps` training steps after `warmup`.  The process group (NCCL) must already be initialised when
    world > 1.  Returns the result dict on every rank (times are the max over ranks)."""
    import torch.distributed as dist
    import densematchingbenchmark_b200 as P
    from densematchingbenchmark_b200 import _cabi
    from densematchingbenchmark_b200.utils.dist_utils import GradReducer, enable_sync_batchnorm, peer_comm
    from densematchingbenchmark_b200.modeling.stereo.backbones.PSMNet import PSMNetBackbone
    from densematchingbenchmark_b200.modeling.stereo.layers.basic_layers import FusedConvUnit
    from densematchingbenchmark_b200.modeling.stereo.losses.stereo_focal_loss import StereoFocalLoss

    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True

    def exchanges():
        c = peer_comm(create=False)
        return c.device_exchanges() if c is not None else 0

    cfg = P.ConfigDict(model=dict(
        batch_norm=True,
        cost_processor=dict(type="Concatenation",
                            cost_computation=dict(type="default", max_disp=max_disp // 4, start_disp=0, dilation=1),
                            cost_aggregator=dict(type=kind, max_disp=max_disp, in_planes=64)),
        disp_predictor=dict(type="FASTER", max_disp=max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True)))
    torch.manual_seed(0)                                             # identical replicas
    bb = PSMNetBackbone(3).to(device).train() if backbone else None
    if bb is not None and channels_last_backbone:
        # cuDNN's fastest 2-D kernels are NHWC: with NCHW tensors every convolution of the (torch) backbone is wrapped
        # in a pair of layout-conversion kernels (~1400 launches, ~10 ms of the step: profiles/r2_launches_train*.csv)
        bb = bb.to(memory_format=torch.channels_last)
    proc = P.build_cost_processor(cfg).to(device).train()
    pred = P.build_disp_predictor(cfg).to(device).train()
    synced = bool(sync_bn and world > 1)
    n_bn = 0
    if synced:
        n_bn = enable_sync_batchnorm(proc)
        if bb is not None and sync_backbone_bn:                      # dmb/apis/train.py:95-97 converts the WHOLE model
            from densematchingbenchmark_b200.utils.dist_utils import convert_sync_batchnorm
            bb = convert_sync_batchnorm(bb)                          # (torch's own SyncBatchNorm synchronises with the
            #                                                          host once per layer: +40 ms on this step)
    params = list(proc.parameters()) + (list(bb.parameters()) if bb is not None else [])
    opt = torch.optim.RMSprop(params, lr=1e-3)                       # configs/PSMNet/scene_flow.py:134
    reducer = GradReducer(params, bucket_mb=bucket_mb) if world > 1 else None
    use_focal = (loss == "config" and kind == "AcfNet")
    focal = StereoFocalLoss(max_disp, 0, 1, weights=(1.0, 0.7, 0.5), focal_coefficient=5.0) if use_focal else None
    l1_weight = 0.1 if use_focal else 1.0

    g = torch.Generator().manual_seed(1000 + rank)                   # different pairs per rank
    B, H, W = batch, height, width
    if bb is not None:
        left = torch.rand(B, 3, H, W, generator=g).to(device)
        right = torch.rand(B, 3, H, W, generator=g).to(device)
        if channels_last_backbone:
            left = left.contiguous(memory_format=torch.channels_last)
            right = right.contiguous(memory_format=torch.channels_last)
    else:
        left = (torch.randn(B, 32, H // 4, W // 4, generator=g) * 0.5).to(device).requires_grad_(True)
        right = (torch.randn(B, 32, H // 4, W // 4, generator=g) * 0.5).to(device).requires_grad_(True)
    gt = (torch.rand(B, 1, H, W, generator=g) * 149 + 1).to(device)
    weights = (1.0, 0.7, 0.5)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    def losses(lf, rf):
        costs = proc(lf, rf)
        disps = [pred(c) for c in costs]
        # mean over the valid pixels without boolean indexing (d[mask] sizes its result on the host: a
        # synchronisation per level, and not capturable): the same sum divided by the same count
        mask = ((gt > 0) & (gt < max_disp)).to(gt.dtype)
        n_valid = mask.sum().clamp_min(1.0)
        total = l1_weight * sum(w * (torch.nn.functional.smooth_l1_loss(d, gt, reduction="none") * mask).sum() / n_valid
                                for w, d in zip(weights, disps))
        if focal is not None:
            total = total + sum(focal(list(costs), gt, variance=1.2).values())
        return total

    class _HotPath(torch.nn.Module):
        """cat volume + aggregator + soft-argmin + the configuration's losses as one module: with `graph_hot_path` its
        forward and backward are one CUDA-graph launch each (torch.cuda.make_graphed_callables)."""

        def __init__(self):
            super(_HotPath, self).__init__()
            self.proc, self.pred = proc, pred

        def forward(self, lf, rf):
            return losses(lf, rf)

    views = hot = None
    if bb is not None and graph_backbone:
        # the torch backbone's two per-view passes as CUDA graphs (forward and backward one launch each): its ~120
        # BatchNorm layers make the eager pass launch bound, and with synchronised statistics every layer is a rendezvous
        from densematchingbenchmark_b200.utils.dist_utils import graph_backbone_views
        if synced and sync_backbone_bn:
            peer_comm()                                              # set up (collectively) before the capture
        if not (synced and sync_backbone_bn) or peer_comm(create=False) is not None:   # (NCCL fall-back: eager)
            views = graph_backbone_views(bb, left, views=2)
    if graph_hot_path and (not synced or peer_comm() is not None):
        with torch.no_grad():
            f0 = (bb._forward(left) if bb is not None else left).detach()
        sample = (f0.clone().requires_grad_(True), f0.clone().requires_grad_(True))
        hot = torch.cuda.make_graphed_callables(_HotPath(), sample, num_warmup_iters=2, allow_unused_input=True)

    def step(marks=None):
        def mark():
            if marks is not None:
                marks.append(ev()); marks[-1].record()
        opt.zero_grad(set_to_none=True)
        mark()
        if views is not None:
            lf, rf = views[0](left), views[1](right)
        else:
            lf, rf = bb(left, right) if bb is not None else (left, right)
        total = hot(lf, rf) if hot is not None else losses(lf, rf)
        mark()
        total.backward()
        mark()
        if reducer is not None:
            reducer.finish()
        torch.nn.utils.clip_grad_norm_(params, 35.0)                 # optimizer_config.grad_clip
        mark()
        opt.step()
        mark()
        return total

    ex0 = exchanges()                                                # (after the graph capture's own warm-up passes)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = _cabi.launch_count()
    all_marks = []
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(steps):
        m = []
        total = step(m)
        all_marks.append(m)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1) / steps
    seg = [sum(m[i].elapsed_time(m[i + 1]) for m in all_marks) / steps for i in range(4)]
    if world > 1:
        t = torch.tensor([ms] + seg, device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, seg = float(t[0]), [float(v) for v in t[1:]]
    grad_bytes = sum(p.numel() * p.element_size() for p in params if p.requires_grad)
    bn_channels = sum(m.bn.num_features for m in proc.modules() if isinstance(m, FusedConvUnit) and m.bn is not None)
    return {
        "workload": "%s training step, %d pairs/GPU of %dx%d, D=%d, fp32 parameters and activations, tcgen05 convolutions "
                    "in split 16-bit arithmetic (config 5)" % (kind, B, H, W, max_disp),
        "n_gpus": world, "pairs_per_gpu": B, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "pairs_per_s": B * world / (ms * 1e-3),
        "segments_ms": {"forward+loss": seg[0], "backward (all-reduce overlapped)": seg[1],
                        "reducer tail + grad clip": seg[2], "optimizer": seg[3]},
        "loss_terms": ("stereo focal loss (coefficient 5, variance 1.2) + 0.1 x smooth-L1" if use_focal else "smooth-L1"),
        "hot_path_launch": "CUDA graphs (forward + losses / backward one launch each)" if hot is not None else "eager",
        "sync_bn": synced, "sync_bn_layers": n_bn,
        "backbone_sync_bn": bool(synced and bb is not None and sync_backbone_bn),
        "backbone": ("torch autograd (cuDNN), outside the hot path"
                     + ("; its two per-view passes replayed from CUDA graphs" if views is not None else ", eager"))
                    if bb is not None else "none (synthetic features)",
        "collective": ("NCCL all-reduce (mean) of all gradients, %d buckets of <= %.0f MB issued from inside backward; "
                       "SyncBN statistics (2*C numbers per BatchNorm layer and direction): %s"
                       % (len(reducer.buckets), bucket_mb,
                          ("one peer-memory kernel per exchange over NVLink P2P stores (csrc/peer_comm.cu), %d exchanges per step"
                           % ((exchanges() - ex0) // max(1, steps + warmup)))
                          if (synced and peer_comm(create=False) is not None) else "NCCL all-reduce per layer"))
                      if reducer is not None else "none (1 GPU)",
        "nccl_bytes_per_step": (int(grad_bytes + (2 * 2 * 8 * bn_channels
                                                   if (synced and peer_comm(create=False) is None) else 0)) if world > 1 else 0),
        "peer_exchanges_per_step": (exchanges() - ex0) // max(1, steps + warmup),
        "buckets_reduced_inside_backward": (reducer.launched_early // max(1, steps + warmup) if reducer is not None else 0),
        "library_launches_per_step": (_cabi.launch_count() - n0) // steps,
        "loss": float(total), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="AcfNet", choices=["AcfNet", "PSMNet"])
    ap.add_argument("--batch", type=int, default=4, help="pairs per GPU (configs/AcfNet/*: 32 over 8 GPUs)")
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--max-disp", type=int, default=192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--no-backbone", action="store_true", help="feed synthetic features (hot path only)")
    ap.add_argument("--bucket-mb", type=float, default=4.0)
    ap.add_argument("--local-backbone-bn", action="store_true", help="keep the torch backbone's BatchNorm per rank")
    ap.add_argument("--nchw-backbone", action="store_true", help="torch backbone in NCHW (default: channels_last)")
    ap.add_argument("--eager-backbone", action="store_true", help="do not capture the backbone's passes in CUDA graphs")
    ap.add_argument("--graph-hot-path", action="store_true", help="capture the hot path's forward + losses / backward in CUDA graphs too")
    ap.add_argument("--loss", default="config", choices=["config", "l1"],
                    help="config: the losses of the reference configuration; l1: smooth-L1 only")
    args = ap.parse_args()

    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        res = run_train_bench(args.kind, args.batch, args.height, args.width, args.max_disp, args.steps, args.warmup,
                              args.sync_bn, not args.no_backbone, args.bucket_mb, args.loss, rank, world, device,
                              not args.local_backbone_bn, not args.nchw_backbone, not args.eager_backbone,
                              args.graph_hot_path)
        if rank == 0:
            print(json.dumps(res))
    finally:
        if world > 1:
            try:
                from densematchingbenchmark_b200.utils.dist_utils import close_peer_comms
                close_peer_comms()
            except Exception:
                pass
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

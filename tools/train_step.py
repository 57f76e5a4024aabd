"""BASELINE config 5: one data-parallel training step of the stereo model (AcfNet / PSMNet aggregator) on
synthetic SceneFlow-sized crops -- backbone (plain torch autograd, outside the hot path) + cat volume +
aggregator (training-mode BatchNorm) + soft-argmin, smooth-L1 loss, backward through the library's own
kernels, bucketed NCCL gradient all-reduce overlapped with backward, RMSprop step.

    python tools/train_step.py                                   # 1 GPU, 4 pairs of 256x512, D=192
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_step.py --sync-bn                            # 8 x 4 pairs

Prints one JSON line (rank 0): ms per step (max over ranks, CUDA events), pairs/s over all ranks, and the
split forward / backward(+overlapped all-reduce) / reducer tail / optimizer.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
    args = ap.parse_args()

    import torch.distributed as dist
    import densematchingbenchmark_b200 as P
    from densematchingbenchmark_b200.utils.dist_utils import GradReducer, enable_sync_batchnorm
    from densematchingbenchmark_b200.modeling.stereo.backbones.PSMNet import PSMNetBackbone

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    torch.backends.cudnn.benchmark = True

    cfg = P.ConfigDict(model=dict(
        batch_norm=True,
        cost_processor=dict(type="Concatenation",
                            cost_computation=dict(type="default", max_disp=args.max_disp // 4, start_disp=0, dilation=1),
                            cost_aggregator=dict(type=args.kind, max_disp=args.max_disp, in_planes=64)),
        disp_predictor=dict(type="FASTER", max_disp=args.max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True)))
    torch.manual_seed(0)                                             # identical replicas
    backbone = None if args.no_backbone else PSMNetBackbone(3).to(device).train()
    proc = P.build_cost_processor(cfg).to(device).train()
    pred = P.build_disp_predictor(cfg).to(device).train()
    if args.sync_bn and world > 1:
        enable_sync_batchnorm(proc)
    params = list(proc.parameters()) + (list(backbone.parameters()) if backbone is not None else [])
    opt = torch.optim.RMSprop(params, lr=1e-3)                       # configs/PSMNet/scene_flow.py:134
    reducer = GradReducer(params, bucket_mb=args.bucket_mb) if world > 1 else None

    g = torch.Generator().manual_seed(1000 + rank)                   # different pairs per rank
    B, H, W = args.batch, args.height, args.width
    if backbone is not None:
        left = torch.rand(B, 3, H, W, generator=g).to(device)
        right = torch.rand(B, 3, H, W, generator=g).to(device)
    else:
        left = (torch.randn(B, 32, H // 4, W // 4, generator=g) * 0.5).to(device).requires_grad_(True)
        right = (torch.randn(B, 32, H // 4, W // 4, generator=g) * 0.5).to(device).requires_grad_(True)
    gt = (torch.rand(B, 1, H, W, generator=g) * 149 + 1).to(device)
    weights = (1.0, 0.7, 0.5)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(marks=None):
        def mark():
            if marks is not None:
                marks.append(ev()); marks[-1].record()
        opt.zero_grad(set_to_none=True)
        mark()
        lf, rf = backbone(left, right) if backbone is not None else (left, right)
        costs = proc(lf, rf)
        disps = [pred(c) for c in costs]
        mask = (gt > 0) & (gt < args.max_disp)
        loss = sum(w * torch.nn.functional.smooth_l1_loss(d[mask], gt[mask]) for w, d in zip(weights, disps))
        mark()
        loss.backward()
        mark()
        if reducer is not None:
            reducer.finish()
        torch.nn.utils.clip_grad_norm_(params, 35.0)                 # optimizer_config.grad_clip
        mark()
        opt.step()
        mark()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from densematchingbenchmark_b200 import _cabi
    n0 = _cabi.launch_count()
    all_marks = []
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        m = []
        loss = step(m)
        all_marks.append(m)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    seg = [sum(m[i].elapsed_time(m[i + 1]) for m in all_marks) / args.steps for i in range(4)]
    if world > 1:
        t = torch.tensor([ms] + seg, device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, seg = float(t[0]), [float(v) for v in t[1:]]
    if rank == 0:
        print(json.dumps({
            "workload": "%s training step, %d pairs/GPU of %dx%d, D=%d, fp32 (config 5)" % (args.kind, B, H, W, args.max_disp),
            "n_gpus": world, "ms_per_step": ms, "pairs_per_s": B * world / (ms * 1e-3),
            "segments_ms": {"forward+loss": seg[0], "backward (all-reduce overlapped)": seg[1],
                            "reducer tail + grad clip": seg[2], "optimizer": seg[3]},
            "sync_bn": bool(args.sync_bn and world > 1), "backbone": "torch autograd" if backbone is not None else "none",
            "buckets_reduced_inside_backward": (reducer.launched_early if reducer is not None else 0),
            "library_launches_per_step": (_cabi.launch_count() - n0) // args.steps,
            "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo tests (CPU) of the multi-rank host logic of bench.py: the path shards on the batch
axis with NO data-path collective; the only cross-rank operation is the max-over-ranks of the timings."""
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = bench.max_over_ranks([10.0 + rank, 5.0 - rank], torch.device("cpu"), world)
        q.put((rank, got, bench.aggregate_throughput(1, world, got[0])))
    finally:
        dist.destroy_process_group()


def test_max_over_ranks_and_aggregate_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, thr in res:
        assert got == [11.0, 5.0]                       # max over ranks, identical on every rank
        assert abs(thr - 2 / 11e-3) < 1e-9              # 2 pairs per 11 ms step


def test_reference_arm_only_rank0_prints():
    """Under torchrun the reference arm runs on rank 0 alone; other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from densematchingbenchmark_b200.utils.dist_utils import GradReducer, all_reduce_grads
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def make():
            torch.manual_seed(0)                                   # identical replicas
            return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16),
                                       torch.nn.ReLU(), torch.nn.Linear(16, 3), torch.nn.Linear(3, 2))
        g = torch.Generator().manual_seed(100 + rank)              # different data per rank
        x = torch.randn(5, 6, generator=g)
        # local gradients, then the exact mean over ranks through plain all_reduce as the expectation
        ref = make()
        ref[5].weight.requires_grad_(False)                        # a frozen parameter is skipped
        ref(x).square().mean().backward()
        want = []
        for p in ref.parameters():
            if p.requires_grad:
                t = p.grad.clone()
                dist.all_reduce(t)
                want.append(t / world)
        # (1) reference-compatible function, coalesced with small buckets and un-coalesced
        for kw in (dict(coalesce=True, bucket_size_mb=-1), dict(coalesce=True, bucket_size_mb=1), dict(coalesce=False)):
            m = make(); m[5].weight.requires_grad_(False)
            m(x).square().mean().backward()
            all_reduce_grads(m, **kw)
            got = [p.grad for p in m.parameters() if p.requires_grad]
            ok1 = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, want))
            if not ok1:
                break
        # (2) overlapped reducer with tiny buckets (several collectives launched inside backward), two steps,
        # the second one leaving the last layer without a gradient on every rank
        m = make(); m[5].weight.requires_grad_(False)
        red = GradReducer(m.parameters(), bucket_mb=0.0005)
        m(x).square().mean().backward()
        early = red.launched_early
        red.finish()
        got = [p.grad for p in m.parameters() if p.requires_grad]
        ok2 = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, want)) and early >= 2
        for p in m.parameters():
            p.grad = None
        h = m[3](m[2](m[1](m[0](x))))                             # stops before the last two layers
        h.square().mean().backward()
        red.finish()
        ok3 = m[4].weight.grad is not None and float(m[4].weight.grad.abs().max()) == 0.0
        t = m[0].weight.grad.clone()
        # (3) round-2 advisor finding: a parameter without gradient on ONE rank only (a data-dependent branch).
        # Rank 1 skips the last two layers, so its early buckets never fill; the reducer issues collectives strictly
        # in bucket order, so both ranks still pair bucket i with bucket i and the means are exact.
        for p in m.parameters():
            p.grad = None
        if rank == 1:
            m[3](m[2](m[1](m[0](x)))).square().mean().backward()
        else:
            m(x).square().mean().backward()
        red.finish()
        ref2 = make(); ref2[5].weight.requires_grad_(False)
        if rank == 1:
            ref2[3](ref2[2](ref2[1](ref2[0](x)))).square().mean().backward()
        else:
            ref2(x).square().mean().backward()
        ok4 = True
        for p, pr in zip(m.parameters(), ref2.parameters()):
            if not pr.requires_grad:
                continue
            w = pr.grad.clone() if pr.grad is not None else torch.zeros_like(pr)
            dist.all_reduce(w)
            ok4 = ok4 and torch.allclose(p.grad, w / world, atol=1e-6)
        q.put((rank, bool(ok1), bool(ok2), bool(ok3 and ok4), len(red.buckets), float(t.abs().sum())))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_and_all_reduce_grads_gloo():
    """The training path's one collective (mean all-reduce of gradients), world size 2 on gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29870 + (os.getpid() % 100)
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert res[0][4] >= 3                                        # the tiny bucket size really split the parameters
    assert abs(res[0][5] - res[1][5]) < 1e-6                     # both ranks hold the same averaged gradient

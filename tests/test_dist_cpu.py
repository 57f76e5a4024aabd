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

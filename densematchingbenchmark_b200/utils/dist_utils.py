"""Gradient exchange of the data-parallel training step -- the ONE collective on the path
(reference: dmb/utils/dist_utils.py:16-47, `all_reduce_grads` = SUM over ranks
then divide by world size, called from DistOptimizerHook.after_train_iter after backward()).

`all_reduce_grads` keeps the reference's signature and semantics.  `GradReducer` is the B200 version of
the same exchange: gradients are packed into a few flat buckets in the order the backward pass produces
them and each bucket's all-reduce (NCCL over NVLink/NVSwitch) is launched the moment its last gradient
lands, so the exchange of the classifier / last-hourglass gradients overlaps the backward kernels of the
earlier layers; `finish()` waits, averages and scatters back.  The reference (MMDistributedDataParallel
plus an explicit all_reduce_grads) reduces every gradient twice; this does it once.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _buckets_by_size(tensors, bucket_bytes):
    """Consecutive runs of `tensors` (same dtype/device) whose sizes add up to <= bucket_bytes."""
    buckets, cur, cur_bytes, cur_key = [], [], 0, None
    for t in tensors:
        key = (t.dtype, t.device)
        nbytes = t.numel() * t.element_size()
        if cur and (key != cur_key or (bucket_bytes > 0 and cur_bytes + nbytes > bucket_bytes)):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(t)
        cur_bytes += nbytes
        cur_key = key
    if cur:
        buckets.append(cur)
    return buckets


def all_reduce_grads(model, coalesce=True, bucket_size_mb=-1, group=None):
    """Mean of every parameter gradient over the ranks, in place (dmb/utils/dist_utils.py:38-47)."""
    grads = [p.grad.data for p in model.parameters() if p.requires_grad and p.grad is not None]
    world = _world(group)
    if world == 1 or not grads:
        return
    if not coalesce:
        for g in grads:
            dist.all_reduce(g.div_(world), group=group)
        return
    bucket_bytes = bucket_size_mb * 1024 * 1024 if bucket_size_mb > 0 else -1
    if bucket_bytes > 0:
        buckets = _buckets_by_size(grads, bucket_bytes)
    else:                                           # one bucket per tensor type, like the reference's default
        by_type = {}
        for g in grads:
            by_type.setdefault((g.dtype, g.device), []).append(g)
        buckets = list(by_type.values())
    for bucket in buckets:
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class GradReducer(object):
    """Bucketed, backward-overlapped mean all-reduce of `params`' gradients.

        reducer = GradReducer(model.parameters(), bucket_mb=8)
        loss.backward()          # buckets are reduced while backward is still running
        reducer.finish()         # p.grad now holds the mean over ranks
    """

    def __init__(self, params, bucket_mb=8.0, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        # gradients become ready roughly in reverse registration order
        order = list(reversed(self.params))
        self.buckets = _buckets_by_size(order, int(bucket_mb * 1024 * 1024))
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device) for b in self.buckets]
        self.slot = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self.slot[id(p)] = (bi, off)
                off += p.numel()
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        self._seen = set()
        self._next = 0                             # first bucket whose all-reduce has not been issued yet
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launched_early = 0                    # buckets whose all-reduce started inside backward()

    def _on_grad(self, p):
        if _world(self.group) == 1 or id(p) in self._seen:
            return
        self._seen.add(id(p))
        bi, off = self.slot[id(p)]
        self.flat[bi][off:off + p.numel()].copy_(p.grad.reshape(-1))
        self._pending[bi] -= 1
        # Collectives are issued STRICTLY in bucket order (bucket i only after buckets 0..i-1): a parameter that
        # gets no gradient on one rank only (a data-dependent branch) then delays that rank's remaining buckets to
        # finish() instead of re-ordering them -- every rank issues the same sequence of all-reduces, whatever
        # order its gradients arrive in (NCCL pairs collectives by issue order).
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._work[self._next] = dist.all_reduce(self.flat[self._next], group=self.group, async_op=True)
            self._next += 1
            self.launched_early += 1

    def finish(self):
        """Complete the exchange: launch what backward() did not (parameters without a gradient this step
        contribute zeros, so that every rank issues the same collectives), wait, average, write back."""
        world = _world(self.group)
        if world > 1:
            for bi, b in enumerate(self.buckets):
                if self._work[bi] is None:
                    for p in b:
                        if id(p) not in self._seen:
                            _, off = self.slot[id(p)]
                            seg = self.flat[bi][off:off + p.numel()]
                            if p.grad is None:
                                seg.zero_()
                            else:
                                seg.copy_(p.grad.reshape(-1))
                    self._work[bi] = dist.all_reduce(self.flat[bi], group=self.group, async_op=True)
            for bi, b in enumerate(self.buckets):
                self._work[bi].wait()
                self.flat[bi].div_(world)
                for p in b:
                    _, off = self.slot[id(p)]
                    seg = self.flat[bi][off:off + p.numel()].view_as(p)
                    if p.grad is None:
                        p.grad = seg.clone()
                    else:
                        p.grad.copy_(seg)
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        self._seen = set()
        self._next = 0

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def enable_sync_batchnorm(module, group=True):
    """Synchronise the batch statistics of every fused conv unit across `group` (True = the default
    group) -- the counterpart of apex.parallel.convert_syncbn_model in dmb/apis/train.py:95-97."""
    from ..modeling.stereo.layers.basic_layers import FusedConvUnit
    n = 0
    for m in module.modules():
        if isinstance(m, FusedConvUnit):
            m.sync_group = group
            n += 1
    return n

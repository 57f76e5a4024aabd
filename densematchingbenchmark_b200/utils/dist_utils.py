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


# ---------------------------------------------------------------------------------------------------------
# Synchronised-BatchNorm statistics over peer memory (csrc/peer_comm.cu) instead of ~300 tiny NCCL collectives per
# step.  NCCL stays the transport of the one real exchange on the path, the gradient all-reduce.
# ---------------------------------------------------------------------------------------------------------
class PeerComm(object):
    """Per-process-group receive buffers shared through CUDA IPC.  `exchange` is one kernel on the current stream.
    All ranks must call it in the same order, from one stream."""
    MAX_BYTES = 8192

    def __init__(self, group=None):
        import ctypes
        from .. import _cabi as C
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("PeerComm covers the GPUs of one box (<= 8 ranks)")
        self._C, self._ct = C, ctypes
        self._own, self._imported, self.seq, self.exchanges = None, [], 0, 0
        self._bufs = (ctypes.c_void_p * self.world)()
        # phase 1 (local, then one collective every rank reaches): allocate, export, gather the handles
        handle = ctypes.create_string_buffer(64)
        err = None
        try:
            own = ctypes.c_void_p()
            C.call("dmb_b200_peer_alloc", ctypes.byref(own))
            self._own = own
            C.call("dmb_b200_peer_export", own, handle)
        except Exception as e:                        # noqa: BLE001
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.rank, bytes(handle.raw) if err is None else None), group=group)
        # phase 2 (local): map the peers' buffers; the caller (peer_comm) agrees on success across ranks before anybody
        # launches an exchange -- a rank that failed here must not leave the others spinning on its flags
        try:
            if err is not None or any(h[1] is None for h in handles):
                raise RuntimeError("a rank could not allocate / export its receive buffer: %s" % (err,))
            for r, raw in sorted(handles):
                if r == self.rank:
                    self._bufs[r] = self._own
                else:
                    ptr = ctypes.c_void_p()
                    C.call("dmb_b200_peer_import", ctypes.create_string_buffer(raw, 64), ctypes.byref(ptr))
                    self._bufs[r] = ptr
                    self._imported.append(ptr)
            self.error = None
        except Exception as e:                        # noqa: BLE001
            self.error = e

    def exchange(self, src, mode):
        """mode 'gather' -> [world, n]; 'sum' -> same shape as src (float64 or float32).  src: contiguous CUDA tensor
        of at most 8192 bytes."""
        C = self._C
        if not src.is_contiguous():
            src = src.contiguous()
        nbytes = src.numel() * src.element_size()
        if nbytes > self.MAX_BYTES or nbytes % 4:
            raise ValueError("PeerComm.exchange: %d bytes (multiple of 4, <= %d)" % (nbytes, self.MAX_BYTES))
        if mode == "gather":
            dst, m = torch.empty((self.world,) + tuple(src.shape), dtype=src.dtype, device=src.device), 0
        elif src.dtype == torch.float64:
            dst, m = torch.empty_like(src), 1
        elif src.dtype == torch.float32:
            dst, m = torch.empty_like(src), 2
        else:
            raise TypeError("PeerComm.exchange('sum') takes float32 / float64")
        self.exchanges += 1                       # (launches issued from Python: a replayed CUDA graph is not counted)
        # sequence number 0: the buffer's own device-side counter, so the launch can sit inside a captured graph
        C.call("dmb_b200_peer_exchange", self._bufs, self.rank, self.world, 0, C.ptr(src), C.ptr(dst), nbytes, m,
               C.stream(src.device))
        return dst

    def sum2(self, a, b):
        """Float32 sums over the ranks of two vectors in ONE exchange -> (sum_a, sum_b)."""
        C = self._C
        a, b = a.contiguous(), b.contiguous()
        if a.dtype != torch.float32 or b.dtype != torch.float32 or (a.numel() + b.numel()) * 4 > self.MAX_BYTES:
            raise ValueError("PeerComm.sum2: two float32 vectors of at most %d bytes together" % self.MAX_BYTES)
        oa, ob = torch.empty_like(a), torch.empty_like(b)
        self.exchanges += 1
        C.call("dmb_b200_peer_sum2_f32", self._bufs, self.rank, self.world, C.ptr(a), a.numel(), C.ptr(b), b.numel(),
               C.ptr(oa), C.ptr(ob), C.stream(a.device))
        return oa, ob

    def bn_forward(self, mean, invstd, count, eps, momentum, running_mean, running_var):
        """Per-rank BatchNorm statistics (torch.batch_norm_stats) -> statistics of the joint batch, running statistics
        updated in place, every rank's element count: exchange and merge in one kernel.  Returns (mean, invstd,
        counts int32 [world])."""
        C = self._C
        mean, invstd = mean.contiguous(), invstd.contiguous()
        om, oi = torch.empty_like(mean), torch.empty_like(invstd)
        counts = torch.empty(self.world, dtype=torch.int32, device=mean.device)
        self.exchanges += 1
        C.call("dmb_b200_peer_bn_forward", self._bufs, self.rank, self.world, C.ptr(mean), C.ptr(invstd), float(count),
               float(eps), float(momentum), C.ptr(running_mean), C.ptr(running_var), C.ptr(om), C.ptr(oi), C.ptr(counts),
               mean.numel(), C.stream(mean.device))
        return om, oi, counts

    def device_exchanges(self):
        """Exchanges this rank has completed, read from the device-side counter (counts the launches replayed by
        CUDA graphs too; synchronises the device)."""
        out = self._ct.c_longlong(0)
        self._C.call("dmb_b200_peer_count", self._own, self._ct.byref(out))
        return int(out.value)

    def close(self, collective=True):
        C = self._C
        torch.cuda.synchronize()
        if collective:
            dist.barrier(group=self.group)
        for ptr in self._imported:
            C.call("dmb_b200_peer_close", ptr)
        self._imported = []
        if self._own is not None:
            C.call("dmb_b200_peer_free", self._own)
            self._own = None


_PEER_COMMS = {}


def peer_comm(group=None, create=True):
    """The PeerComm of `group` (None = default group), created on first use; None when peer memory is unavailable
    (single GPU, no CUDA IPC / P2P between the ranks' devices, DMB_B200_PEER_COMM=0): callers then use NCCL."""
    import os
    key = id(group) if group not in (None, True) else None
    if key in _PEER_COMMS:
        return _PEER_COMMS[key]
    if not create or os.environ.get("DMB_B200_PEER_COMM", "1") == "0" or not dist.is_initialized():
        return None
    g = None if group in (None, True) else group
    if not torch.cuda.is_available() or not (1 < dist.get_world_size(g) <= 8) or dist.get_backend(g) != "nccl":
        _PEER_COMMS[key] = None                       # (gloo / single rank: nothing to set up, the same on every rank)
        return None
    comm = PeerComm(g)                                # collective inside: every rank gets here
    ok = torch.tensor([0 if comm.error is not None else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=g)      # all ranks must agree: one rank falling back alone would
    if int(ok.item()) == 0:                                 # leave the others spinning on its flags
        import warnings
        warnings.warn("dmb_b200: peer-memory exchange unavailable (%s); SyncBN statistics go through NCCL" % (comm.error,))
        comm.close(collective=False)
        comm = None
    else:
        dist.barrier(group=g)                         # every buffer is zeroed and mapped before the first exchange
        probe = comm.exchange(torch.full((4,), float(comm.rank + 1), device="cuda", dtype=torch.float64), "sum")
        if not bool((probe == sum(range(1, comm.world + 1))).all()):
            raise RuntimeError("dmb_b200: peer-memory exchange self-test failed")
    _PEER_COMMS[key] = comm
    return comm


def close_peer_comms():
    for comm in _PEER_COMMS.values():
        if comm is not None:
            comm.close()
    _PEER_COMMS.clear()


def sum_over_ranks_(t, group=None):
    """In-place SUM of a small statistics tensor over the ranks: peer-memory kernel when available, else NCCL."""
    comm = peer_comm(group)
    if comm is not None and t.numel() * t.element_size() <= PeerComm.MAX_BYTES and t.dtype in (torch.float32, torch.float64):
        t.copy_(comm.exchange(t, "sum"))
    else:
        dist.all_reduce(t, group=None if group in (None, True) else group)
    return t


class _PeerSyncBatchNormFn(torch.autograd.Function):
    """Synchronised BatchNorm of a torch module (the 2-D backbone) with torch's own statistics / normalisation kernels
    (batch_norm_stats, batch_norm_gather_stats_with_counts, batch_norm_elemt and their backward twins) and the
    exchange of the per-rank statistics through `PeerComm` -- the arithmetic of torch.nn.SyncBatchNorm without its
    all_gather collective and without the host synchronisation that its empty-rank mask costs per layer."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group):
        if not (x.is_contiguous(memory_format=torch.channels_last) or x.is_contiguous()):
            x = x.contiguous()
        C_ = x.shape[1]
        mean, invstd = torch.batch_norm_stats(x, eps)
        comm = peer_comm(group)
        fused = (comm is not None and mean.dtype == torch.float32 and (2 * C_ + 1) * 4 <= PeerComm.MAX_BYTES
                 and all(t is None or (t.dtype == torch.float32 and t.is_contiguous()) for t in (running_mean, running_var)))
        if fused:
            # exchange + merge + running-statistics update in one kernel (csrc/peer_comm.cu:peer_bn_forward_kernel)
            mean, invstd, counts = comm.bn_forward(mean, invstd, x.numel() // C_, eps, momentum, running_mean, running_var)
        else:
            count = torch.full((1,), x.numel() // C_, dtype=mean.dtype, device=mean.device)
            combined = torch.cat([mean, invstd, count], dim=0)                       # [2C + 1]
            if comm is not None:
                allc = comm.exchange(combined, "gather")                             # [world, 2C + 1]
            else:
                world = dist.get_world_size(None if group in (None, True) else group)
                allc = torch.empty(world, combined.numel(), dtype=combined.dtype, device=combined.device)
                dist.all_gather_into_tensor(allc.view(-1), combined, None if group in (None, True) else group)
            mean_all, invstd_all, count_all = torch.split(allc, C_, dim=1)
            counts = count_all.reshape(-1)
            mean, invstd = torch.batch_norm_gather_stats_with_counts(x, mean_all, invstd_all, running_mean, running_var,
                                                                     momentum, eps, counts)
            counts = counts.to(torch.int32)
        ctx.save_for_backward(x, weight, mean, invstd, counts)
        ctx.group = group
        return torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(ctx, gy):
        if not (gy.is_contiguous(memory_format=torch.channels_last) or gy.is_contiguous()):
            gy = gy.contiguous()
        x, weight, mean, invstd, counts = ctx.saved_tensors
        sum_dy, sum_dy_xmu, gw, gb = torch.batch_norm_backward_reduce(gy, x, mean, invstd, weight, ctx.needs_input_grad[0],
                                                                      ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        gx = None
        if ctx.needs_input_grad[0]:
            comm = peer_comm(ctx.group)
            if comm is not None and sum_dy.dtype == torch.float32 and 2 * sum_dy.numel() * 4 <= PeerComm.MAX_BYTES:
                sum_dy, sum_dy_xmu = comm.sum2(sum_dy, sum_dy_xmu)                # both sums in one exchange, no cat / split
            else:
                combined = torch.cat([sum_dy, sum_dy_xmu], dim=0)
                sum_over_ranks_(combined, ctx.group)
                sum_dy, sum_dy_xmu = torch.split(combined, sum_dy.shape[0])
            gx = torch.batch_norm_backward_elemt(gy, x, mean, invstd, weight, sum_dy, sum_dy_xmu, counts)
        return gx, (gw if ctx.needs_input_grad[1] else None), (gb if ctx.needs_input_grad[2] else None), None, None, None, None, None


class PeerSyncBatchNorm(torch.nn.modules.batchnorm._BatchNorm):
    """BatchNorm over (N, C, ...) whose batch statistics are taken over all ranks of `group`; parameters, buffers and
    state-dict keys of the BatchNorm it replaces (counterpart of apex.parallel.SyncBatchNorm for the torch parts of
    the model, dmb/apis/train.py:95-97)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, group=None):
        super(PeerSyncBatchNorm, self).__init__(num_features, eps, momentum, affine, track_running_stats)
        self.group = group

    def _check_input_dim(self, x):
        if x.dim() < 2:
            raise ValueError("expected at least 2D input (got {}D input)".format(x.dim()))

    def forward(self, x):
        world = dist.get_world_size(None if self.group in (None, True) else self.group) if dist.is_initialized() else 1
        if not self.training or world == 1 or not x.is_cuda:
            return torch.nn.functional.batch_norm(x, self.running_mean if not self.training or self.track_running_stats else None,
                                                  self.running_var if not self.training or self.track_running_stats else None,
                                                  self.weight, self.bias, self.training, self.momentum or 0.0, self.eps)
        momentum = self.momentum
        if self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(self.num_batches_tracked)
        return _PeerSyncBatchNormFn.apply(x, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
                                          self.running_var if self.track_running_stats else None, self.eps,
                                          momentum if momentum is not None else 0.0, self.group)


def convert_sync_batchnorm(module, group=None):
    """Replace every torch BatchNorm{1,2,3}d below `module` by a PeerSyncBatchNorm sharing its parameters and buffers
    (what apex.parallel.convert_syncbn_model / nn.SyncBatchNorm.convert_sync_batchnorm do in the reference's trainer)."""
    out = module
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm) and not isinstance(module, PeerSyncBatchNorm):
        out = PeerSyncBatchNorm(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats, group)
        if module.affine:
            out.weight, out.bias = module.weight, module.bias
        out.running_mean, out.running_var, out.num_batches_tracked = module.running_mean, module.running_var, module.num_batches_tracked
        out.training = module.training
    for name, child in module.named_children():
        new = convert_sync_batchnorm(child, group)
        if new is not child:
            out.add_module(name, new)
    return out


class _BackboneView(torch.nn.Module):
    """One per-view pass of a backbone (`backbone._forward`) as a module of its own: the unit `graph_backbone_views`
    captures.  The views share the backbone's parameters and buffers."""

    def __init__(self, backbone):
        super(_BackboneView, self).__init__()
        self.backbone = backbone

    def forward(self, x):
        return self.backbone._forward(x)


def graph_backbone_views(backbone, sample, views=2, warmup=3):
    """CUDA graphs of the per-view training passes of a torch backbone (forward and backward of each view one graph
    launch each, torch.cuda.make_graphed_callables).  The ~120 BatchNorm layers of the PSMNet backbone make its eager
    training pass launch bound; synchronised over ranks (PeerSyncBatchNorm) every layer is in addition a rendezvous,
    so host-side launch jitter of ANY rank stalls all of them ~240 times per step.  Captured, the exchanges (device-
    side sequence counter, csrc/peer_comm.cu) follow each other at GPU speed.  Every rank must call this at the same
    point (the warm-up passes exchange statistics).  Returns `views` callables image -> features, to be called in
    order within a step; BatchNorm running statistics are updated by the replays as in eager mode."""
    if not sample.is_cuda:
        raise ValueError("graph_backbone_views: CUDA tensors only")
    mods = tuple(_BackboneView(backbone) for _ in range(views))
    args = tuple((sample.detach().clone(memory_format=torch.preserve_format),) for _ in range(views))
    return torch.cuda.make_graphed_callables(mods, args, num_warmup_iters=warmup)


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

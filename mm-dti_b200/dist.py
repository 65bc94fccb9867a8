"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed; NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The reference is single-process (SURVEY.md §2:
"no distributed code"); this is the new work §8(e) names:

  exchange 1 (forward)   all-gather of the contrastive operands (unit embeddings of both
                         modalities, pooled features, labels / predictions / weights) so that
                         InfoNCE / SupCon / ConR see global-batch negatives; a second, tiny
                         all-gather of the per-row statistics before the gradient phase
  exchange 2 (backward)  gradient all-reduce (bucketed, flat buffers)
  FDS epoch statistics   all-reduce of per-bucket {count, sum} and {second moment}

Each rank evaluates ITS anchor rows against the global keys and obtains the complete gradient
of the global loss w.r.t. its own rows, so no reduce-scatter of key-side gradients is needed
(ops_sim.py).  Ranks must hold equal local batch sizes."""
import torch
import torch.distributed as dist


class DataParallelCtx:
    """Handle passed as ``dp=`` to the contrastive ops / FDS.

    grad_scale: factor applied to the local gradients of a GLOBAL loss.  Every rank returns the
    same global loss and the exact d(global loss)/d(local rows); a gradient all-reduce that
    AVERAGES over ranks (DistributedDataParallel, ``allreduce_grads(average=True)``) would divide
    it by world, so the default grad_scale = world cancels that; use 1.0 with a summing reduce."""

    def __init__(self, group=None, grad_scale=None):
        if not dist.is_initialized():
            raise RuntimeError("DataParallelCtx needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.grad_scale = float(self.world if grad_scale is None else grad_scale)

    def all_gather_rows(self, t):
        """(M, ...) on every rank -> (world*M, ...), rank-major."""
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def all_reduce_sum(self, t):
        t = t.contiguous().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t


def allreduce_grads(params, group=None, average=True, bucket_bytes=64 << 20, async_op=True):
    """Exchange 2: all-reduce ``p.grad`` of every parameter in flat same-dtype buckets of about
    ``bucket_bytes`` (NVSwitch: size buckets for launch latency, not link count).  Parameters whose
    grad is None on this rank contribute zeros (every rank must issue the same collectives).
    async_op=False issues plain in-stream collectives (use it inside a CUDA-graph capture)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    by_dtype = {}
    for p in params:
        if not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        by_dtype.setdefault(p.grad.dtype, []).append(p.grad)
    n_coll = 0
    for dt, grads in by_dtype.items():
        bucket, size = [], 0
        buckets = []
        for g in grads:
            nb = g.numel() * g.element_size()
            if bucket and size + nb > bucket_bytes:
                buckets.append(bucket)
                bucket, size = [], 0
            bucket.append(g)
            size += nb
        if bucket:
            buckets.append(bucket)
        works = []
        for b in buckets:
            flat = torch.cat([g.reshape(-1) for g in b])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op), flat, b))
            n_coll += 1
        for w, flat, b in works:
            if async_op:
                w.wait()
            if average:
                flat.div_(world)
            off = 0
            for g in b:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
    return n_coll


class OverlappedGradReducer:
    """Exchange 2 overlapped with the backward pass: parameters are grouped into buckets of about ``bucket_bytes``
    in REVERSE registration order (the order their gradients become ready); a post-accumulate-grad hook counts a
    bucket down and, when it is complete, all-reduces it on a communication stream while the main stream keeps
    running the rest of the backward.  ``finish()`` (call it after ``loss.backward()``) reduces whatever is left
    (parameters that received no gradient contribute zeros) and joins the streams.  Every rank issues the same
    collectives in the same order.  Works inside a CUDA-graph capture (the communication stream becomes a parallel
    branch of the graph); on CPU tensors (gloo tests) the reduction runs synchronously.

    tail_params: parameters whose gradients arrive LAST although they are registered early (here: the token embedding
    and the K1 pair-bias parameters, whose backward runs after layer 0).  They get a bucket of their own at the end, so
    the bucket holding the first encoder layers is reduced while K1's backward still runs and only a few KB stay exposed.

    keep_flat=True: every bucket owns a persistent flat buffer (16-byte aligned slots); a complete bucket is gathered
    into it with one multi-tensor copy and all-reduced IN PLACE -- no division, no copy back.  ``p.grad`` then keeps the
    LOCAL gradient; the reduced one is ``reduced_grad(p)`` (a view of the flat buffer), to be scaled by ``grad_scale``
    (1/world when average=True).  ``optim.FusedAdam(grad_source=reducer.reduced_grad, grad_scale=reducer.grad_scale)``
    consumes it directly."""

    def __init__(self, params, group=None, average=True, bucket_bytes=32 << 20, tail_params=None, keep_flat=False,
                 comm_dtype=None):
        """comm_dtype (keep_flat mode): dtype the gradients travel in, e.g. torch.bfloat16 -- the bucket is gathered into a
        bf16 wire buffer (one multi-tensor copy, which also converts), all-reduced there (half the NVLink bytes and half
        the time the NCCL kernels share the SMs with the backward), and widened back into the fp32 flat buffer that
        ``reduced_grad`` exposes.  ``bucket_bytes`` counts fp32 parameter bytes either way."""
        self.group, self.average, self.keep_flat = group, average, keep_flat
        self.comm_dtype = comm_dtype if keep_flat else None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        tail = [p for p in (tail_params or []) if p.requires_grad]
        tail_ids = {id(p) for p in tail}
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            if id(p) in tail_ids:
                continue
            nb = p.numel() * p.element_size()
            if cur and (size + nb > bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nb
        if cur:
            self.buckets.append(cur)
        for p in tail:                                   # same-dtype runs of the late parameters
            if self.buckets and getattr(self.buckets[-1], "is_tail", False) and self.buckets[-1][0].dtype == p.dtype:
                self.buckets[-1].append(p)
            else:
                b = _Bucket([p])
                b.is_tail = True
                self.buckets.append(b)
        self.bucket_of = {p: i for i, b in enumerate(self.buckets) for p in b}
        self.comm = torch.cuda.Stream(device=self.params[0].device) if self.params[0].is_cuda else None
        self.flat, self.slot, self.wire, self.wire_slot = [], {}, [], {}
        if keep_flat:
            for b in self.buckets:
                offs, off = [], 0
                for p in b:
                    offs.append(off)
                    off += (p.numel() + 3) // 4 * 4
                flat = torch.zeros(off, device=b[0].device, dtype=b[0].dtype)
                self.flat.append(flat)
                for p, o in zip(b, offs):
                    self.slot[p] = flat[o:o + p.numel()].view_as(p)
                if self.comm_dtype is not None and self.comm_dtype != b[0].dtype:
                    wire = torch.zeros(off, device=b[0].device, dtype=self.comm_dtype)
                    self.wire.append(wire)
                    for p, o in zip(b, offs):
                        self.wire_slot[p] = wire[o:o + p.numel()].view_as(p)
                else:
                    self.wire.append(None)
        self._reset()
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.collectives = 0

    @property
    def grad_scale(self):
        return 1.0 / self.world if self.average else 1.0

    def reduced_grad(self, p):
        """keep_flat mode: the all-reduced (SUMMED) gradient of p after finish(); multiply by grad_scale."""
        return self.slot[p]

    def _reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.done = [False] * len(self.buckets)

    def _hook(self, p):
        if self.world == 1:
            return
        i = self.bucket_of[p]
        self.pending[i] -= 1
        if self.pending[i] == 0 and not self.done[i]:
            self._reduce(i)

    def _reduce(self, i):
        bucket = self.buckets[i]
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p)

        def run():
            if self.keep_flat:
                if self.wire[i] is not None:
                    torch._foreach_copy_([self.wire_slot[p] for p in bucket], [p.grad for p in bucket])      # gather + narrow
                    dist.all_reduce(self.wire[i], op=dist.ReduceOp.SUM, group=self.group)
                    self.flat[i].copy_(self.wire[i])                                                          # widen
                    return
                torch._foreach_copy_([self.slot[p] for p in bucket], [p.grad for p in bucket])
                dist.all_reduce(self.flat[i], op=dist.ReduceOp.SUM, group=self.group)
                return
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                flat.div_(self.world)
            off = 0
            for p in bucket:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n

        if self.comm is not None:
            main = torch.cuda.current_stream(bucket[0].device)
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                run()
        else:
            run()
        self.done[i] = True
        self.collectives += 1

    def finish(self):
        if self.world > 1:
            for i in range(len(self.buckets)):
                if not self.done[i]:
                    self._reduce(i)
            if self.comm is not None:
                torch.cuda.current_stream(self.params[0].device).wait_stream(self.comm)
        self._reset()

    def remove(self):
        for h in self.handles:
            h.remove()


class _Bucket(list):
    """list of parameters with room for an attribute"""
    is_tail = False


def shard_rows(n_total, rank, world):
    """[start, stop) of the rows a rank owns when n_total samples are split evenly."""
    if n_total % world:
        raise ValueError("global batch %d is not divisible by the world size %d" % (n_total, world))
    m = n_total // world
    return rank * m, rank * m + m

"""Data-parallel use of the dense per-anchor path: images shard across ranks, nothing else moves.

Every stage is independent per image (reference anchors.py:393, bench.py:44,69); the only
cross-image coupling is the loss normaliser sum(num_positives)+1 (loss.py:261) and the final sums.
So the sharded path needs two latency-bound collectives (the shapes the reference's own
``effdet/distributed.py`` helpers would move, evaluator.py:38-39):
  * all-reduce(sum) of the scalar normaliser before the loss kernel (gradients are scaled by the
    GLOBAL 1/N inside the same pass), then all-reduce(sum) of the three loss scalars;
  * all-gather of the padded detections [B_local, D, 6] (+ counts, + OOD scores).
With these the result equals the single-process reference at the global batch, up to fp32
summation order.  Works with any initialised ``torch.distributed`` backend (NCCL over NVLink on
the B200 box, gloo in the CPU tests).
"""
from typing import List, Optional

import torch
import torch.distributed as dist


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_range(global_batch: int, rank: int, world_size: int):
    """Images [lo, hi) owned by ``rank`` (contiguous blocks, remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_normalizer(num_positives: torch.Tensor, group=None) -> torch.Tensor:
    """sum over ALL ranks' images of num_positives, + 1 (loss.py:261) -> fp32 [1] on every rank."""
    local = num_positives.float().sum().reshape(1)
    if _active(group):
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local + 1.0


def reduce_losses(total, cls_loss, box_loss, group=None):
    """Per-rank partial losses (each already divided by the global normaliser) -> global values."""
    packed = torch.stack([total.detach().reshape(()), cls_loss.detach().reshape(()), box_loss.detach().reshape(())])
    if _active(group):
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0], packed[1], packed[2]


class PendingLosses:
    """Result of ``forward_losses_one_collective(..., async_op=True)``: the all-reduce runs on NCCL's own
    stream while the caller keeps enqueueing the next step; ``result()`` waits and normalises."""

    def __init__(self, packed, work, box_loss_weight):
        self.packed, self.work, self.w = packed, work, box_loss_weight

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        n = self.packed[2] + 1.0
        cls_loss, box_loss = self.packed[0] / n, self.packed[1] / n
        return cls_loss + self.w * box_loss, cls_loss, box_loss


def forward_losses_one_collective(cls_unnorm, box_unnorm, num_positives, box_loss_weight, group=None, async_op=False):
    """Forward-only variant with ONE collective: every rank computes its partial sums
    against a unit normaliser, a single all-reduce carries [sum_cls, sum_box, sum_num_positives], and
    the division by the global (num_positives + 1) happens afterwards.  Returns (total, cls, box).
    (With gradients use ``sharded_detection_loss``: the backward scale needs the global N up front.)"""
    packed = torch.stack([cls_unnorm.detach().reshape(()).float(), box_unnorm.detach().reshape(()).float(),
                          num_positives.float().sum().reshape(())])
    work = None
    if _active(group):
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    pending = PendingLosses(packed, work if async_op else None, box_loss_weight)
    return pending if async_op else pending.result()


class PendingPartialSums:
    """Result of ``all_reduce_partial_sums(..., async_op=True)``; ``result()`` waits and normalises."""

    def __init__(self, packed, work, world):
        self.packed, self.work, self.world = packed, work, world

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        # slot 3 is sum_r (sum(num_positives_r) + 1): the global normaliser is that minus (world - 1)
        res = self.packed[:3] / (self.packed[3] - float(self.world - 1))
        return res[0], res[1], res[2]


def local_partial_sums(labeler, cls_outputs, box_outputs, gt_boxes, gt_classes, unit, buf=None, **loss_kw):
    """This rank's share of the forward loss with no torch kernel in between: the labeler writes
    sum(num_positives) + 1 into slot 3 of one 4-float buffer (``buf``, allocated when None) and the
    fused loss (against the unit normaliser ``unit``) writes [cls + w * box, cls, box] partial sums
    into slots 0..2."""
    from .loss import loss_fn_fused
    if buf is None:
        buf = torch.empty((4,), dtype=torch.float32, device=gt_boxes.device)
    lb = labeler.assign(gt_boxes, gt_classes, normalizer_out=buf[3:4])
    loss_fn_fused(cls_outputs, box_outputs, lb, normalizer=unit, out=buf, **loss_kw)
    return buf


def all_reduce_partial_sums(buf, group=None, async_op=False, copy=True):
    """ONE all-reduce of the 4 floats of ``local_partial_sums``; returns (total, cls, box) of the GLOBAL
    batch (or a ``PendingPartialSums``).  With ``copy`` the buffer is cloned first, so the caller may
    overwrite it right away (the next CUDA-graph replay does) while the collective is still in flight;
    ``copy=False`` reduces in place (see ``LossReducePipeline`` for the ordering that makes that safe)."""
    packed = buf.detach().clone() if copy else buf.detach()
    world, work = 1, None
    if _active(group):
        world = dist.get_world_size(group)
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    pending = PendingPartialSums(packed, work if async_op else None, world)
    return pending if async_op else pending.result()


class LossReducePipeline:
    """Keeps the per-step collective and its bookkeeping OFF the compute stream.

    ``submit(buf)`` (called on the compute stream right after the kernels that fill ``buf``) records an
    event and, on a side stream, waits for it, all-reduces ``buf`` in place and normalises.  The compute
    stream never waits for the collective: it only has to wait for the returned ``done`` event before
    the same ``buf`` is written again (two alternating buffers / CUDA graphs make that wait free).
    ``collect()`` returns the oldest (total, cls, box) -- tensors produced on the side stream -- and its
    ``done`` event; synchronise on it (or the device) before reading them elsewhere."""

    def __init__(self, device, group=None):
        self.group = group
        self.side = torch.cuda.Stream(device=device)
        self.pending = []

    def submit(self, buf):
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            res = all_reduce_partial_sums(buf, self.group, async_op=True, copy=False).result()
            done = torch.cuda.Event()
            done.record(self.side)
        self.pending.append((res, done))
        return done

    def collect(self):
        return self.pending.pop(0)

    def __len__(self):
        return len(self.pending)


def sharded_detection_loss(loss_module, cls_outputs, box_outputs, label_batch, group=None):
    """Local shard's fused loss against the GLOBAL normaliser.

    Returns ((total, cls, box) local partials that carry the autograd graph -- their sum over
    ranks is the global loss, so ``total.backward()`` on every rank yields exactly the gradients
    of the global loss w.r.t. the local outputs -- and the all-reduced (total, cls, box) for logging)."""
    norm = global_normalizer(label_batch.num_positives, group)
    part = loss_module.forward_fused(cls_outputs, box_outputs, label_batch, normalizer=norm)
    return part, reduce_losses(*part, group=group)


def gather_detections(dets: torch.Tensor, count: torch.Tensor, extras: Optional[List[torch.Tensor]] = None, group=None):
    """all-gather of the padded per-image results in rank order (equal B_local on every rank).

    dets [B_local, D, 6], count [B_local]; extras: further [B_local, ...] tensors (e.g. OOD scores).
    Returns the same list of tensors with leading dimension B_local * world_size."""
    tensors = [dets, count] + list(extras or [])
    if not _active(group):
        return tensors
    world = dist.get_world_size(group)
    out = []
    for t in tensors:
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        out.append(torch.cat(parts, dim=0))
    return out

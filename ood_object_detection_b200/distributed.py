"""Data-parallel use of the dense per-anchor path: images shard across ranks, nothing else moves.

Every stage is independent per image (reference anchors.py:393, bench.py:44,69); the only
cross-image coupling is the loss normaliser sum(num_positives)+1 (loss.py:261) and the final sums
(what the reference's own ``effdet/distributed.py`` helpers would move, evaluator.py:38-39):
  * forward / logging path: every rank computes partial sums against a unit normaliser and the 4
    floats [cls + w*box, cls, box, sum(num_pos)+1] are exchanged ONCE per step -- by remote stores into
    peer mailboxes over NVLink issued by the loss kernel itself (``PeerMailbox``,
    ``local_partial_sums(..., mailbox=)``), or by one all-reduce (``all_reduce_partial_sums``);
  * gradient path: all-reduce(sum) of the scalar normaliser BEFORE the loss kernel (gradients are scaled
    by the GLOBAL 1/N inside the same pass), then all-reduce(sum) of the three loss scalars
    (``sharded_detection_loss``);
  * all-gather of the padded detections [B_local, D, 6] (+ counts, + OOD scores).
With these the result equals the single-process reference at the global batch, up to fp32
summation order.  The collective helpers work with any initialised ``torch.distributed`` backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests); the mailboxes need CUDA peer access.
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


def local_partial_sums(labeler, cls_outputs, box_outputs, gt_boxes, gt_classes, unit, buf=None, mailbox=None, **loss_kw):
    """This rank's share of the forward loss with no torch kernel in between: the labeler writes
    sum(num_positives) + 1 into slot 3 of one 4-float buffer (``buf``, allocated when None) and the
    fused loss (against the unit normaliser ``unit``) writes [cls + w * box, cls, box] partial sums
    into slots 0..2.  With ``mailbox`` (a ``PeerMailbox``) the loss kernel itself trades the four floats
    with the other ranks: no collective, no further launch."""
    from .loss import loss_fn_fused
    if buf is None:
        buf = torch.empty((4,), dtype=torch.float32, device=gt_boxes.device)
    lb = labeler.assign(gt_boxes, gt_classes, normalizer_out=buf[3:4], transient=True)
    loss_fn_fused(cls_outputs, box_outputs, lb, normalizer=unit, out=buf,
                  exchange=None if mailbox is None else mailbox.attach(buf[3:4]), **loss_kw)
    return buf


def all_reduce_partial_sums(buf, group=None, async_op=False, copy=True):
    """ONE all-reduce of the 4 floats of ``local_partial_sums``; returns (total, cls, box) of the GLOBAL
    batch (or a ``PendingPartialSums``).  With ``copy`` the buffer is cloned first, so the caller may
    overwrite it right away (the next CUDA-graph replay does) while the collective is still in flight;
    ``copy=False`` reduces in place (the caller orders the next write after the collective, as bench.py's
    alternating CUDA graphs do)."""
    packed = buf.detach().clone() if copy else buf.detach()
    world, work = 1, None
    if _active(group):
        world = dist.get_world_size(group)
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    pending = PendingPartialSums(packed, work if async_op else None, world)
    return pending if async_op else pending.result()


class PeerMailbox:
    """The loss partial sums over NVLink peer memory instead of a collective (include/odk.h,
    ``odk_partials_publish`` / ``odk_partials_collect``): ``publish(buf4)`` stores this rank's 4 floats into
    every rank's mailbox and returns at once; ``collect()`` (one step later, every rank) sums the records
    in rank order and returns (total, cls, box) of the global batch plus a status flag.  Both are single
    tiny kernels with device-side sequence counters, so they can be captured in a CUDA graph.  No
    collective kernel sits on an SM while the persistent loss grid runs, and ranks are coupled with one
    step of slack instead of in lockstep.

    The mailboxes are torch symmetric memory (CUDA VMM handles exchanged through the process group's
    store); ``world == 1`` (or ``local_only``) uses plain device memory.  Raises when symmetric memory
    cannot be set up -- callers fall back to ``all_reduce_partial_sums`` (NCCL)."""

    def __init__(self, device, group=None, local_only=False, timeout_ms=30000):
        from . import _lib
        self._lib = _lib
        lib = _lib.lib()
        self.device = torch.device(device)
        active = _active(group) and not local_only
        self.world = dist.get_world_size(group) if active else 1
        self.rank = dist.get_rank(group) if active else 0
        nbytes = int(lib.odk_mailbox_bytes(self.world))
        if nbytes == 0:
            raise ValueError(f'world size {self.world} not supported by the mailbox exchange')
        if active:
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else dist.group.WORLD
            enable = getattr(symm, 'enable_symm_mem_for_group', None)
            if enable is not None:      # needed by older torch, a deprecated no-op in newer ones
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    try:
                        enable(grp.group_name)
                    except Exception:
                        pass
            self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.buf.zero_()
            self.handle = symm.rendezvous(self.buf, grp)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier(group)          # every mailbox is zero before anybody publishes
        else:
            self.buf = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)
            self.handle = None
            ptrs = [self.buf.data_ptr()]
        self._ptrs = (_lib.ctypes.c_void_p * self.world)(*ptrs)
        self._ptr_list = ptrs
        self.published = self.collected = 0
        self.timeout_ms = int(timeout_ms)   # how long a collect waits for the peers (wall clock); the status flag is sticky
        self.out3 = torch.zeros((3,), dtype=torch.float32, device=self.device)
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)

    def attach(self, num_pos_plus_1, normalized=False):
        """Descriptor for ``loss_fn_fused(..., exchange=...)``: the loss kernel's finishing CTA collects the
        previous step's records into ``self.out3`` / ``self.status`` (when one is outstanding) and publishes
        its own sums together with ``num_pos_plus_1`` (float32 [1], what the labeler wrote)."""
        _lib = self._lib
        if num_pos_plus_1.dtype != torch.float32 or num_pos_plus_1.numel() != 1 or num_pos_plus_1.device != self.device:
            raise ValueError('num_pos_plus_1 must be one float32 element on the mailbox device')
        x = _lib.Exchange()
        for r, p in enumerate(self._ptr_list):
            x.mailboxes[r] = p
        x.world, x.rank = self.world, self.rank
        x.num_pos_plus_1 = num_pos_plus_1.data_ptr()
        x.global_out3, x.status = self.out3.data_ptr(), self.status.data_ptr()
        x.normalized, x.timeout_ms = int(bool(normalized)), int(self.timeout_ms)
        self._keep = (x, num_pos_plus_1)
        return x

    def previous(self):
        """(total, cls, box) of the last step a fused launch or ``collect()`` has collected (NaN if that step's
        records did not arrive in time; ``check()`` raises then)."""
        return self.out3[0], self.out3[1], self.out3[2]

    def check(self):
        """Raise if any collect so far timed out (the device flag is sticky: 1 + the first failing sequence number).
        Reads one int from the device: call it where a sync is acceptable (end of an epoch, logging step)."""
        st = int(self.status.item())
        if st != 0:
            raise RuntimeError(f'peer mailbox exchange: the records of step {st - 1} did not arrive within '
                               f'{self.timeout_ms} ms (a rank is gone or stalled); loss values since then are NaN')

    def publish(self, buf4):
        if buf4.dtype != torch.float32 or buf4.numel() < 4 or not buf4.is_contiguous() or buf4.device != self.device:
            raise ValueError('publish needs a contiguous float32 [4] tensor on the mailbox device')
        _lib = self._lib
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().odk_partials_publish(_lib.ptr(buf4), self._ptrs, self.world, self.rank,
                                                       _lib.stream_ptr(self.device)))
        self.published += 1

    def collect(self, out=None, status=None):
        """-> ((total, cls, box) views of ``out``, status int32 [1]: 0 ok, 1 a peer's record never arrived)."""
        _lib = self._lib
        out = self.out3 if out is None else out
        status = self.status if status is None else status
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().odk_partials_collect(_lib.ptr(self.buf), self.world, _lib.ptr(out), _lib.ptr(status),
                                                       self.timeout_ms, _lib.stream_ptr(self.device)))
        self.collected += 1
        return (out[0], out[1], out[2]), status


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
        whole = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        try:      # one collective straight into the result (rank order), capturable in a CUDA graph
            dist.all_gather_into_tensor(whole, t, group=group)
        except (RuntimeError, NotImplementedError, AttributeError):   # a backend without the flat form
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t, group=group)
            whole = torch.cat(parts, dim=0)
        out.append(whole)
    return out

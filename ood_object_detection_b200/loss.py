"""Detection loss with the reference's API (effdet/loss.py), computed by libodk's fused kernel.

``loss_fn`` / ``DetectionLoss`` (reference :224-298, :355-401) keep their signatures and return
``(total, class_loss, box_loss)`` scalars that support ``.backward()``.  One odk_loss launch
streams the NCHW logits once, never builds the one-hot tensor, and (when a gradient is needed)
writes d total / d outputs in the same pass.  ``loss_fn_fused`` additionally skips the target
tensors altogether by consuming the labeler's ``match`` directly.

The fork's meta-learning extras that are not on the dense hot path (``SupportLoss``,
``smooth_l1_loss``, ``l2_loss``, ``cosine_loss``, ``huber_loss`` and the two element-wise focal
functions) are plain tensor code with the reference's semantics.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


# ------------------------------------------------------------------------------- fused path
def _as_f32c(t):
    t = _lib.require_cuda(t, 'loss input')
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _one_block(tensors, dtype):
    """Reference-layout per-level targets -> the single concatenated block odk_loss reads.
    Zero-copy when the tensors are consecutive views of one buffer (what AnchorLabeler returns)."""
    first = tensors[0]
    esz = first.element_size()
    nxt = first.data_ptr()
    consecutive = True
    for t in tensors:
        if t.dtype != dtype or not t.is_contiguous() or t.data_ptr() != nxt:
            consecutive = False
            break
        nxt += t.numel() * esz
    if consecutive and first.data_ptr() % 16 == 0:
        return first, list(tensors)
    return torch.cat([t.to(dtype).reshape(-1) for t in tensors]), list(tensors)


class _DetectionLossFn(torch.autograd.Function):
    """autograd bridge: forward = one odk_loss launch (gradients of the total are produced in
    the same pass when any input requires grad); backward only rescales the stored gradients
    by the upstream factors (a no-op kernel when they are 1, as in ``loss.backward()``)."""

    @staticmethod
    def forward(ctx, meta, *outputs):
        n = meta['levels']
        layout = 0
        if meta.get('label_batch') is not None:
            # fused targets: channels_last head outputs (AMP / channels_last models) are read in place
            cls_out, cmask = _lib.prep_levels(outputs[:n], n, 'loss input')
            box_out, bmask = _lib.prep_levels(outputs[n:2 * n], n, 'loss input')
            layout = cmask | (bmask << 8)
        else:   # targets given as tensors (the reference's layout): the plane-walking kernel, NCHW only
            cls_out = [_as_f32c(t) for t in outputs[:n]]
            box_out = [_as_f32c(t) for t in outputs[n:2 * n]]
        dev = cls_out[0].device
        lib = _lib.lib()
        need_grad = any(ctx.needs_input_grad[1:])
        ctx.set_materialize_grads(False)   # unused outputs arrive as None in backward: no zero-fill kernels
        B = cls_out[0].shape[0]
        C = meta['num_classes']
        hw = [c.shape[2] * c.shape[3] for c in cls_out]
        na = box_out[0].shape[1] // 4
        for c, b in zip(cls_out, box_out):
            if c.shape[1] != na * C or b.shape[0] != B or c.shape[0] != B:
                raise ValueError(f'class outputs {tuple(c.shape)} do not match num_classes={C}, anchors/location={na}')
        out = meta.get('out')
        if out is None:
            out = torch.empty((3,), dtype=torch.float32, device=dev)
        elif out.dtype != torch.float32 or out.numel() < 3 or not out.is_contiguous() or out.device != dev:
            raise ValueError('out= must be a contiguous float32 tensor with >= 3 elements on the outputs\' device')
        ws = torch.empty((lib.odk_loss_workspace_bytes() + 15) // 16 * 2, dtype=torch.int64, device=dev)
        gcls = gbox = None
        gcls_p = gbox_p = None
        if need_grad:
            # one owning tensor per level: autograd can adopt them as .grad without a copy
            gcls = [torch.empty_like(t) for t in cls_out]
            gbox = [torch.empty_like(t) for t in box_out]
            gcls_p, gbox_p = _lib.ptr_array(gcls), _lib.ptr_array(gbox)
        fused = meta.get('label_batch')
        if fused is not None and fused.consumed:
            raise RuntimeError('this LabelBatch was created with transient=True and has already been consumed by a loss call')
        use_keys = fused is not None and fused.keys is not None
        clear = use_keys and fused.transient and fused.workspace is not None
        exchange = meta.get('exchange')   # distributed.PeerMailbox.attach(...) descriptor, or None
        params = _lib.LossParams(meta['alpha'], meta['gamma'], meta['delta'], meta['box_loss_weight'],
                                 meta['label_smoothing'], int(bool(meta['legacy_focal'])), int(use_keys), int(clear), layout,
                                 _lib.ctypes.pointer(exchange) if exchange is not None else None)
        if fused is not None:
            match = fused.keys if use_keys else fused.match
            anchors, gtb, gtl = fused.labeler._anchor_tables(dev)[0], fused.gt_boxes, fused.gt_labels
            cls_t = box_t = None
            mmax = gtb.shape[1]
        else:
            match = anchors = gtb = gtl = None
            cls_t, box_t, mmax = meta['cls_block'], meta['box_block'], 0
        with torch.cuda.device(dev):
            _lib.check(lib.odk_loss(_lib.ptr_array(cls_out), _lib.ptr_array(box_out), B, C, _lib.int_array(hw), n, na,
                                    _lib.ptr(match), _lib.ptr(anchors), _lib.ptr(gtb), _lib.ptr(gtl), mmax,
                                    _lib.ptr(cls_t), _lib.ptr(box_t), _lib.ptr(meta['normalizer']), params,
                                    _lib.ptr(out), gcls_p, gbox_p, _lib.ptr(ws), ws.numel() * 8, _lib.stream_ptr(dev)))
        if clear:   # the kernel's last CTA zeroed the keys: the workspace goes back to the labeler, the batch is spent
            fused.consumed = True
            fused.labeler._recycle(fused.workspace)
            fused.keys = fused.workspace = None
        ctx.gcls, ctx.gbox = (gcls, gbox) if need_grad else (None, None)
        ctx.box_loss_weight = meta['box_loss_weight']
        ctx.levels = n
        ctx.in_dtypes = [t.dtype for t in outputs]
        total, cls_loss, box_loss = out[0], out[1], out[2]
        return total, cls_loss, box_loss

    @staticmethod
    def backward(ctx, g_total, g_cls, g_box):
        lib = _lib.lib()
        if ctx.gcls is None:
            # the buffers were scaled in place and handed to autograd by the first backward (no copy, no second
            # 1.2 GB pass); a second walk over the graph would need them back
            raise RuntimeError('the fused detection loss supports ONE backward pass per forward (its gradient buffers are '
                               'handed to autograd as .grad): call loss_fn again instead of retain_graph=True, and '
                               'backpropagate a single combination of (loss, class_loss, box_loss)')
        dev = ctx.gcls[0].device
        # stored: d total / d logits (= d cls_loss / d logits) and d total / d box (= w * d box_loss / d box);
        # in the usual total.backward() both factors ARE g_total and no torch kernel runs here
        w = ctx.box_loss_weight
        if w == 0 and g_box is not None:
            raise RuntimeError('box_loss_weight == 0: the stored box gradient is w * d box_loss / d box = 0, so a gradient '
                               'through box_loss alone cannot be recovered; use a non-zero weight or detach box_loss')

        def factor(a, b):
            if a is None and b is None:
                return torch.zeros((1,), dtype=torch.float32, device=dev)
            t = a if b is None else (b if a is None else a.float() + b.float())
            return t.float().reshape(1).contiguous()

        s_cls = factor(g_total, g_cls)
        s_box = factor(g_total, None if g_box is None else g_box.float() / w)
        # total.backward(): one factor for everything, one launch
        groups = (((ctx.gcls + ctx.gbox), s_cls),) if (g_cls is None and g_box is None) else ((ctx.gcls, s_cls), (ctx.gbox, s_box))
        with torch.cuda.device(dev):
            for bufs, sc in groups:
                for lo in range(0, len(bufs), 16):
                    part = bufs[lo:lo + 16]
                    sizes = (_lib.c_int64 * len(part))(*[t.numel() for t in part])
                    _lib.check(lib.odk_scale_inplace_multi(_lib.ptr_array(part), sizes, len(part), _lib.ptr(sc),
                                                           _lib.stream_ptr(dev)))
        # hand the buffers over (no reference kept here) so autograd can adopt them as .grad without a copy
        bufs = ctx.gcls + ctx.gbox
        ctx.gcls = ctx.gbox = None
        grads = [g.to(dt) if dt != torch.float32 else g for g, dt in zip(bufs, ctx.in_dtypes)]
        del bufs
        return (None, *[g if need else None for g, need in zip(grads, ctx.needs_input_grad[1:])])


def _normalizer(num_positives):
    # loss.py:261: sum of positives over the batch, +1 so an all-background batch is finite
    return (num_positives.float().sum() + 1.0).reshape(1)


def loss_fn(
        cls_outputs: List[torch.Tensor],
        box_outputs: List[torch.Tensor],
        cls_targets: List[torch.Tensor],
        box_targets: List[torch.Tensor],
        num_positives: torch.Tensor,
        num_classes: int,
        alpha: float,
        gamma: float,
        delta: float,
        box_loss_weight: float,
        label_smoothing: float = 0.,
        legacy_focal: bool = False,
        normalizer: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Total detection loss over all levels (reference effdet/loss.py:224-298).

    cls_outputs[l] [B, na*C, H, W], box_outputs[l] [B, na*4, H, W] (NCHW head outputs);
    cls_targets[l] [B, H, W, na] int64 (-1 background, -2 ignored); box_targets[l] [B, H, W, na*4].
    ``normalizer`` (extension): a precomputed global sum(num_positives)+1, used by the sharded path.
    """
    levels = len(cls_outputs)
    dev = cls_outputs[0].device
    cls_block, _keep1 = _one_block([_lib.require_cuda(t, 'cls_targets') for t in cls_targets], torch.int64)
    box_block, _keep2 = _one_block([_lib.require_cuda(t, 'box_targets') for t in box_targets], torch.float32)
    if normalizer is None:
        normalizer = _normalizer(num_positives.to(dev))
    meta = dict(levels=levels, num_classes=int(num_classes), alpha=float(alpha), gamma=float(gamma), delta=float(delta),
                box_loss_weight=float(box_loss_weight), label_smoothing=float(label_smoothing),
                legacy_focal=bool(legacy_focal), cls_block=cls_block, box_block=box_block,
                normalizer=normalizer.float().reshape(1).contiguous())
    return _DetectionLossFn.apply(meta, *cls_outputs, *box_outputs)


def loss_fn_fused(cls_outputs, box_outputs, label_batch, num_classes: int, alpha: float, gamma: float, delta: float,
                  box_loss_weight: float, label_smoothing: float = 0., legacy_focal: bool = False,
                  normalizer: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, exchange=None):
    """Same value as ``loss_fn(.., *labeler.batch_label_anchors(..))`` without ever materialising
    the target tensors: ``label_batch`` is ``AnchorLabeler.assign(...)``.  ``out`` (float32, >= 3
    elements) receives [total, cls_loss, box_loss] in place; the returned scalars are views of it.
    ``exchange``: ``distributed.PeerMailbox.attach(...)`` -- the kernel also trades the partial sums with
    the other data-parallel ranks (use with a unit ``normalizer``)."""
    if normalizer is None:
        normalizer = label_batch.normalizer if label_batch.normalizer is not None else _normalizer(label_batch.num_positives)
    meta = dict(levels=len(cls_outputs), num_classes=int(num_classes), alpha=float(alpha), gamma=float(gamma),
                delta=float(delta), box_loss_weight=float(box_loss_weight), label_smoothing=float(label_smoothing),
                legacy_focal=bool(legacy_focal), label_batch=label_batch,
                normalizer=normalizer.float().reshape(1).contiguous(), out=out, exchange=exchange)
    return _DetectionLossFn.apply(meta, *cls_outputs, *box_outputs)


class DetectionLoss(nn.Module):
    """Reference effdet/loss.py:355-401."""

    __constants__ = ['num_classes']

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.num_classes = config.num_classes
        self.alpha = config.alpha
        self.gamma = config.gamma
        self.delta = config.delta
        self.box_loss_weight = config.box_loss_weight
        self.label_smoothing = config.label_smoothing
        self.legacy_focal = config.legacy_focal
        self.use_jit = config.jit_loss

    def _kw(self):
        return dict(num_classes=self.num_classes, alpha=self.alpha, gamma=self.gamma, delta=self.delta,
                    box_loss_weight=self.box_loss_weight, label_smoothing=self.label_smoothing,
                    legacy_focal=self.legacy_focal)

    def box_loss(self, box_outputs, box_targets, num_positives):
        return box_only_loss(box_outputs, box_targets, num_positives, alpha=self.alpha, gamma=self.gamma,
                             delta=self.delta, box_loss_weight=self.box_loss_weight,
                             label_smoothing=self.label_smoothing, legacy_focal=self.legacy_focal)

    def forward(self, cls_outputs, box_outputs, cls_targets, box_targets, num_positives):
        return loss_fn(cls_outputs, box_outputs, cls_targets, box_targets, num_positives, **self._kw())

    def forward_fused(self, cls_outputs, box_outputs, label_batch, normalizer=None):
        return loss_fn_fused(cls_outputs, box_outputs, label_batch, normalizer=normalizer, **self._kw())


# ------------------------------------------------------------------------------- tensor extras
def focal_loss_legacy(logits, targets, alpha: float, gamma: float, normalizer):
    """Element-wise legacy focal loss (reference effdet/loss.py:15-47)."""
    ce = F.binary_cross_entropy_with_logits(logits, targets.to(logits.dtype), reduction='none')
    neg = -1.0 * logits
    modulator = torch.exp(gamma * targets * neg - gamma * torch.log1p(torch.exp(neg)))
    loss = modulator * ce
    return torch.where(targets == 1.0, alpha * loss, (1.0 - alpha) * loss) / normalizer


def new_focal_loss(logits, targets, alpha: float, gamma: float, normalizer, label_smoothing: float = 0.01,
                   loss_func=F.binary_cross_entropy_with_logits):
    """Element-wise "new" focal loss as the fork runs it (reference effdet/loss.py:49-95): the
    modulating factor is disabled there, so this is alpha-weighted BCE (gamma unused)."""
    targets = targets.to(logits.dtype)
    weight = None
    if alpha is not None:
        weight = targets * alpha + (1. - targets) * (1. - alpha)
    if label_smoothing > 0.:
        targets = targets * (1. - label_smoothing) + .5 * label_smoothing
    loss = loss_func(logits, targets, reduction='none')
    return (1 / normalizer) * loss if weight is None else (1 / normalizer) * weight * loss


def cosine_loss(input, target, margin=0., reduction='mean'):
    loss = torch.where(target == 1., 1 - input, input - margin)
    return loss.clamp(min=0.).mean()


def huber_loss(input, target, delta: float = 1., weights: Optional[torch.Tensor] = None, size_average: bool = True):
    abs_err = (input - target).abs()
    quadratic = torch.clamp(abs_err, max=delta)
    loss = 0.5 * quadratic.pow(2) + delta * (abs_err - quadratic)
    if weights is not None:
        loss = loss * weights
    return loss.mean() if size_average else loss.sum()


def _signed_weight_sums(err, weights):
    ws = torch.sign(err) * weights
    return ws[ws > 0.].sum(), ws[ws < 0.].sum()


def smooth_l1_loss(input, target, beta: float = 1. / 9, weights: Optional[torch.Tensor] = None,
                   size_average: bool = False):
    """Reference effdet/loss.py:121-154 (returns the fork's extra signed-weight sums)."""
    err = input - target
    abs_err = torch.abs(err)
    loss = abs_err if beta < 1e-5 else torch.where(abs_err < beta, 0.5 * abs_err.pow(2) / beta, abs_err - 0.5 * beta)
    pos = neg = None
    if weights is not None:
        loss = loss * weights
        pos, neg = _signed_weight_sums(err, weights)
    if size_average:
        return loss.mean()
    return loss.sum(), pos, neg


def l2_loss(input, target, beta: float = 1. / 9, weights: Optional[torch.Tensor] = None, size_average: bool = False):
    err = input - target
    loss = err ** 2
    pos = neg = None
    if weights is not None:
        loss = loss * weights
        pos, neg = _signed_weight_sums(err, weights)
    return loss.mean(), pos, neg


def _box_loss(box_outputs, box_targets, num_positives, delta: float = 0.1):
    mask = box_targets != 0.0
    return huber_loss(box_outputs, box_targets, weights=mask, delta=delta, size_average=False) / (num_positives * 4.0)


def one_hot(x, num_classes: int):
    nonneg = (x >= 0).unsqueeze(-1)
    oh = torch.zeros(x.shape + (num_classes,), device=x.device, dtype=torch.float32)
    return oh.scatter(-1, x.unsqueeze(-1) * nonneg, 1) * nonneg


def class_loss_fn(cls_outputs, cls_targets, num_positives, num_classes: int, alpha: float, gamma: float,
                  label_smoothing: float = 0., legacy_focal: bool = False, loss_func=F.binary_cross_entropy_with_logits):
    """The fork's support-set class loss (reference effdet/loss.py:188-221): dense NCHW targets."""
    normalizer = num_positives.sum() + 1.0
    total = []
    for out, tgt in zip(cls_outputs, cls_targets):
        loss = new_focal_loss(out.permute(0, 2, 3, 1), tgt.permute(0, 2, 3, 1), alpha=alpha, gamma=gamma,
                              normalizer=normalizer, label_smoothing=label_smoothing, loss_func=loss_func)
        total.append(loss.sum())
    return torch.stack(total, dim=-1).sum(dim=-1)


def box_only_loss(box_outputs, box_targets, num_positives, alpha: float, gamma: float, delta: float,
                  box_loss_weight: float, label_smoothing: float = 0., legacy_focal: bool = False):
    """Reference effdet/loss.py:303-352."""
    normalizer = num_positives.sum() + 1.0
    parts = [_box_loss(o.permute(0, 2, 3, 1), t, normalizer, delta=delta) for o, t in zip(box_outputs, box_targets)]
    return box_loss_weight * torch.stack(parts, dim=-1).sum(dim=-1)


class SupportLoss(nn.Module):
    """Reference effdet/loss.py:404-439."""

    __constants__ = ['num_classes']

    def __init__(self, config, loss_type):
        super().__init__()
        self.config = config
        self.num_classes = config.num_classes
        self.alpha = config.alpha
        self.gamma = config.gamma
        self.label_smoothing = config.label_smoothing
        self.legacy_focal = config.legacy_focal
        self.use_jit = config.jit_loss
        if loss_type == 'ce':
            self.loss_func = F.binary_cross_entropy_with_logits
        elif loss_type == 'mse':
            self.loss_func = F.mse_loss

    def forward(self, cls_outputs, cls_targets, num_positives, alpha):
        return class_loss_fn(cls_outputs, cls_targets, num_positives, num_classes=self.num_classes, alpha=alpha,
                             gamma=self.gamma, label_smoothing=self.label_smoothing, legacy_focal=self.legacy_focal,
                             loss_func=self.loss_func)

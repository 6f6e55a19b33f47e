"""Anchors, AnchorLabeler and detection generation with the reference's API, running on libodk.

Mirrors effdet/anchors.py of the reference (same names, argument meaning and results):
``Anchors`` (:191-302), ``AnchorLabeler`` (:305-438), ``decode_box_outputs`` (:51-85),
``clip_boxes_xyxy`` (:88-92), ``generate_detections`` (:95-172), ``get_feat_sizes`` (:175-188).
The per-image python loops and the torch/torchvision op chains behind them are replaced by the
sm_100a kernels of libodk.so (odk_assign, odk_targets, odk_detect).
"""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .object_detection import ArgMaxMatcher, FasterRcnnBoxCoder, IouSimilarity, TargetAssigner

MIN_CLASS_SCORE = -5.0
_DUMMY_DETECTION_SCORE = -1e5


def get_feat_sizes(image_size: Tuple[int, int], max_level: int):
    """Feature-map (H, W) for levels 0..max_level; each level halves, rounding up."""
    sizes = [tuple(image_size)]
    for _ in range(max_level):
        h, w = sizes[-1]
        sizes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return sizes


class Anchors(nn.Module):
    """Multiscale RetinaNet anchors (reference effdet/anchors.py:191-302).

    ``boxes`` is the [A, 4] yxyx fp32 table in the reference order
    level -> y -> x -> (scale octave, aspect); it is computed on the host in float64 and rounded
    to fp32 once, exactly as the reference does, because the labeler's IoU thresholds are
    sensitive to the last bit.
    """

    def __init__(self, min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size: Tuple[int, int]):
        super().__init__()
        self.min_level = min_level
        self.max_level = max_level
        self.num_scales = num_scales
        self.aspect_ratios = aspect_ratios
        n_levels = max_level - min_level + 1
        if isinstance(anchor_scale, Sequence):
            assert len(anchor_scale) == n_levels
            self.anchor_scales = anchor_scale
        else:
            self.anchor_scales = [anchor_scale] * n_levels
        assert isinstance(image_size, Sequence) and len(image_size) == 2
        assert image_size[0] % 2 ** max_level == 0, 'Image size must be divisible by 2 ** max_level (128)'
        assert image_size[1] % 2 ** max_level == 0, 'Image size must be divisible by 2 ** max_level (128)'
        self.image_size = tuple(image_size)
        self.feat_sizes = get_feat_sizes(image_size, max_level)
        self.config = self._generate_configs()
        self.register_buffer('boxes', self._generate_boxes())
        # not persistent: the reference's Anchors has only `boxes` in its state_dict, and checkpoints must stay interchangeable
        self.register_buffer('plane_desc', self._generate_plane_desc(), persistent=False)
        self.register_buffer('plane_gen', self._generate_plane_gen(), persistent=False)

    @classmethod
    def from_config(cls, config, img_size=None, min_level=0):
        size = config.image_size if img_size is None else (img_size, img_size)
        return cls(config.min_level + min_level, config.max_level, config.num_scales, config.aspect_ratios,
                   config.anchor_scale, size)

    def _generate_configs(self):
        """{level: [(stride_yx, octave_scale, aspect, anchor_scale), ...]} like the reference."""
        base = self.feat_sizes[0]
        cfg = {}
        for level in range(self.min_level, self.max_level + 1):
            fs = self.feat_sizes[level]
            stride = (base[0] // fs[0], base[1] // fs[1])
            scale = self.anchor_scales[level - self.min_level]
            cfg[level] = [(stride, octave / float(self.num_scales), aspect, scale)
                          for octave in range(self.num_scales) for aspect in self.aspect_ratios]
        return cfg

    def _generate_boxes(self):
        per_level = []
        for level_cfg in self.config.values():
            shapes = []
            for stride, octave_scale, aspect, anchor_scale in level_cfg:
                size_x = anchor_scale * stride[1] * 2 ** octave_scale
                size_y = anchor_scale * stride[0] * 2 ** octave_scale
                if isinstance(aspect, Sequence):
                    ax, ay = aspect[0], aspect[1]
                else:
                    ax = np.sqrt(aspect)
                    ay = 1.0 / ax
                half_x, half_y = size_x * ax / 2.0, size_y * ay / 2.0
                cx = np.arange(stride[1] / 2, self.image_size[1], stride[1])
                cy = np.arange(stride[0] / 2, self.image_size[0], stride[0])
                gx, gy = np.meshgrid(cx, cy)
                gx, gy = gx.ravel(), gy.ravel()
                shapes.append(np.stack([gy - half_y, gx - half_x, gy + half_y, gx + half_x], axis=1))
            # [HW, shapes, 4] -> rows ordered (y, x, shape)
            per_level.append(np.stack(shapes, axis=1).reshape(-1, 4))
        return torch.from_numpy(np.concatenate(per_level, axis=0)).float()

    def get_anchors_per_location(self):
        return self.num_scales * len(self.aspect_ratios)

    def _generate_plane_gen(self):
        """[levels*na, 6] float64: cy0, cx0, sy, sx, half_y, half_x of every (level, shape) grid -- the numbers
        _generate_boxes builds the table from, so the labeler kernel can recompute an anchor instead of gathering it."""
        rows = []
        for level_cfg in self.config.values():
            for stride, octave_scale, aspect, anchor_scale in level_cfg:
                size_x = anchor_scale * stride[1] * 2 ** octave_scale
                size_y = anchor_scale * stride[0] * 2 ** octave_scale
                if isinstance(aspect, Sequence):
                    ax, ay = aspect[0], aspect[1]
                else:
                    ax = np.sqrt(aspect)
                    ay = 1.0 / ax
                rows.append([stride[0] / 2, stride[1] / 2, float(stride[0]), float(stride[1]), size_y * ay / 2.0, size_x * ax / 2.0])
        return torch.tensor(rows, dtype=torch.float64)

    def _generate_plane_desc(self):
        """[levels*na, 12] fp32 description of every (level, shape) anchor grid for odk_assign_grid:
        cy0, cx0, sy, sx, hy, hx, area, W, H, off_l, shape, level -- read off the anchor table itself."""
        boxes = self.boxes.double().numpy()
        na = self.get_anchors_per_location()
        rows, off = [], 0
        for li, level in enumerate(range(self.min_level, self.max_level + 1)):
            h, w = self.feat_sizes[level]
            for a in range(na):
                b00 = boxes[off + a]
                cy0, cx0 = (b00[0] + b00[2]) / 2, (b00[1] + b00[3]) / 2
                sy = (boxes[off + w * na + a][0] + boxes[off + w * na + a][2]) / 2 - cy0 if h > 1 else 1.0
                sx = (boxes[off + na + a][1] + boxes[off + na + a][3]) / 2 - cx0 if w > 1 else 1.0
                hy, hx = (b00[2] - b00[0]) / 2, (b00[3] - b00[1]) / 2
                rows.append([cy0, cx0, sy, sx, hy, hx, 4 * hy * hx, w, h, off, a, li])
            off += h * w * na
        return torch.tensor(rows, dtype=torch.float32)

    # ---- libodk geometry -------------------------------------------------------------------
    def level_hw(self):
        return [self.feat_sizes[l][0] * self.feat_sizes[l][1] for l in range(self.min_level, self.max_level + 1)]

    def level_offsets(self):
        na, off, out = self.get_anchors_per_location(), 0, []
        for hw in self.level_hw():
            out.append(off)
            off += hw * na
        out.append(off)
        return out


class LabelBatch:
    """Device-resident result of the assignment kernels for one batch.

    ``match`` ([B, Apad] int32, planar order) is all the fused loss needs; the reference-layout
    target tensors are only materialised on demand (``targets()``)."""

    def __init__(self, labeler, gt_boxes, gt_labels, match, num_positives, keys=None, normalizer=None, workspace=None,
                 transient=False):
        self.labeler = labeler
        self.workspace = workspace        # the labeler's workspace the keys live in (gt-centric kernel)
        self.transient = transient        # consumed by ONE fused loss, which zeroes the keys again (no memset next time)
        self.consumed = False
        self.gt_boxes = gt_boxes
        self.gt_labels = gt_labels
        self._match = match
        self.keys = keys                  # [B, Apad] 64-bit assignment keys (gt-centric kernel) or None
        self.normalizer = normalizer      # [1] sum(num_positives) + 1, produced by the kernel, or None
        self.num_positives = num_positives
        self._targets = None

    @property
    def match(self):
        """[B, Apad] int32 gt row per anchor (planar order), converted from the keys on first use."""
        if self._match is None:
            if self.consumed:
                raise RuntimeError('this LabelBatch was created with transient=True and its keys were consumed by the loss')
            lib = _lib.lib()
            B, apad = self.keys.shape
            self._match = torch.empty((B, apad), dtype=torch.int32, device=self.keys.device)
            with torch.cuda.device(self.keys.device):
                _lib.check(lib.odk_keys_to_match(_lib.ptr(self.keys), B, self.labeler.anchors.boxes.shape[0],
                                                 _lib.ptr(self._match), _lib.stream_ptr(self.keys.device)))
        return self._match

    def targets(self):
        if self._targets is None:
            self._targets = self.labeler._materialize(self)
        return self._targets


class AnchorLabeler(object):
    """Labeler for multiscale anchor boxes (reference effdet/anchors.py:305-438).

    Inputs must live on (or are moved to) the CUDA device holding ``anchors.boxes``; the
    reference's per-image python loop is one odk_assign + one odk_targets launch sequence.
    """

    def __init__(self, anchors, num_classes: int, match_threshold: float = 0.5):
        similarity_calc = IouSimilarity()
        matcher = ArgMaxMatcher(match_threshold, unmatched_threshold=match_threshold,
                                negatives_lower_than_unmatched=True, force_match_for_each_row=True)
        box_coder = FasterRcnnBoxCoder()
        self.target_assigner = TargetAssigner(similarity_calc, matcher, box_coder)
        self.anchors = anchors
        self.match_threshold = match_threshold
        self.num_classes = num_classes
        self.indices_cache = {}
        self._dev_tables = {}
        self.host_results = False
        self._clean_ws = {}   # (device, bytes) -> workspaces the fused loss has left all-zero again (transient batches)
        # gt-centric kernel (odk_assign_grid) for the regular pyramid grids; the dense kernel
        # (odk_assign) is kept for arbitrary anchor sets and non-positive thresholds
        self.use_grid_kernel = True
        # recompute anchors from the float64 generator (Anchors.plane_gen) instead of gathering them from the table:
        # bit-identical (odk_anchor_table), measured equally fast (the kernel is a chain of short phases, not gather-bound)
        self.use_anchor_generator = False

    # ---- helpers ---------------------------------------------------------------------------
    def _device(self):
        """The CUDA device the kernels run on.  With ``anchors`` on a CUDA device: that one, and results stay
        there.  With ``anchors`` still on the host -- how the reference's datasets build their labeler
        (preloader.py:60-62, dataloader.py:60-66) -- the current CUDA device is used with cached device copies of
        the anchor table, and the targets are copied back to the host so the calling script sees what the
        reference gave it (``host_results``).  There is no CPU implementation."""
        dev = self.anchors.boxes.device
        if dev.type == 'cuda':
            self.host_results = False
            return dev
        if not torch.cuda.is_available():
            raise RuntimeError('AnchorLabeler needs a CUDA device: the assignment only exists as sm_100a kernels (no CPU '
                               'fallback).  In a forked DataLoader worker label in the main process instead '
                               '(ood_object_detection_b200.pipeline.label_batch_targets)')
        self.host_results = True
        return torch.device('cuda', torch.cuda.current_device())

    def _anchor_tables(self, dev):
        """(boxes, plane_desc, plane_gen) on ``dev`` (cached copies when the module itself lives on the host)."""
        if self.anchors.boxes.device == dev:
            return self.anchors.boxes, getattr(self.anchors, 'plane_desc', None), getattr(self.anchors, 'plane_gen', None)
        key = (dev, self.anchors.boxes.data_ptr())
        if self._dev_tables.get('key') != key:
            desc, gen = getattr(self.anchors, 'plane_desc', None), getattr(self.anchors, 'plane_gen', None)
            self._dev_tables = {'key': key, 'boxes': self.anchors.boxes.to(dev), 'desc': None if desc is None else desc.to(dev),
                                'gen': None if gen is None else gen.to(dev)}
        return self._dev_tables['boxes'], self._dev_tables['desc'], self._dev_tables['gen']

    def _pack(self, gt_boxes, gt_classes, filter_valid):
        """-> gt_boxes [B,M,4] fp32, labels [B,M] int32, count [B] int32 or None, on device.  Ragged lists (the
        reference scripts' call form, preloader.py:146, dataloader.py:207-210) are padded with ONE vectorised
        scatter and reach the device in one copy per array -- no per-image launches or copies."""
        dev = self._device()
        if isinstance(gt_boxes, torch.Tensor) and gt_boxes.dim() == 3:
            boxes = gt_boxes.to(dev, torch.float32).contiguous()
            cls = gt_classes if isinstance(gt_classes, torch.Tensor) else torch.stack(list(gt_classes))
            labels = cls.to(dev).reshape(boxes.shape[0], -1).to(torch.int32).contiguous()
            return boxes, labels, None
        lens = [int(b.shape[0]) for b in gt_boxes]
        B, M = len(lens), max([1] + lens)
        for b, n in zip(gt_boxes, lens):
            if n and b.dtype != torch.float32:
                raise ValueError('Invalid tensor type: should be tf.float32')  # BoxList contract
        total = sum(lens)
        rows = np.repeat(np.arange(B), lens)
        cols = np.arange(total) - np.repeat(np.cumsum([0] + lens[:-1]), lens)
        count = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        if total == 0:
            return (torch.zeros((B, M, 4), dtype=torch.float32, device=dev),
                    torch.full((B, M), -1, dtype=torch.int32, device=dev), count)
        flat_b = torch.cat([b.reshape(-1, 4) for b, n in zip(gt_boxes, lens) if n])
        flat_c = torch.cat([c.reshape(-1) for c, n in zip(gt_classes, lens) if n]).to(torch.int32)
        if flat_b.device.type == 'cpu':     # pad on the host (pinned), one H2D per array
            boxes_h = torch.zeros((B, M, 4), dtype=torch.float32).pin_memory()
            labels_h = torch.full((B, M), -1, dtype=torch.int32).pin_memory()
            boxes_h[rows, cols] = flat_b
            labels_h[rows, cols] = flat_c
            return boxes_h.to(dev, non_blocking=True), labels_h.to(dev, non_blocking=True), count
        r, c = torch.from_numpy(rows).to(dev, non_blocking=True), torch.from_numpy(cols).to(dev, non_blocking=True)
        boxes = torch.zeros((B, M, 4), dtype=torch.float32, device=dev)
        labels = torch.full((B, M), -1, dtype=torch.int32, device=dev)
        boxes[r, c] = flat_b.to(dev)
        labels[r, c] = flat_c.to(dev)
        return boxes, labels, count

    @staticmethod
    def _relabel_task_cls(boxes, labels, count, task_cls):
        """Reference anchors.py:396-403 for the whole batch at once, on the device: every gt box that a box of the
        task class overlaps with IoU > 0.9 takes the task class.  boxes [B,M,4], labels [B,M] (padded), count [B]
        or None.  Same fp32 operation order as IouSimilarity (region_similarity_calculator.py:48-73), one
        [B,M,M] pass, no host synchronisation.  (An image with other boxes but none of the task class makes the
        reference raise from ``max`` over an empty dimension; here it is simply left unchanged.)"""
        B, M = labels.shape
        valid = torch.ones_like(labels, dtype=torch.bool) if count is None else \
            torch.arange(M, device=labels.device)[None, :] < count[:, None]
        task = (labels == int(task_cls)) & valid
        ymin, xmin, ymax, xmax = boxes.unbind(-1)
        area = (ymax - ymin) * (xmax - xmin)
        h = (torch.min(ymax[:, :, None], ymax[:, None, :]) - torch.max(ymin[:, :, None], ymin[:, None, :])).clamp(min=0)
        w = (torch.min(xmax[:, :, None], xmax[:, None, :]) - torch.max(xmin[:, :, None], xmin[:, None, :])).clamp(min=0)
        inter = h * w
        union = area[:, :, None] + area[:, None, :] - inter
        iou = torch.where(inter == 0.0, torch.zeros_like(inter), inter / union)
        hit = ((iou > 0.9) & task[:, :, None]).any(1) & valid
        return torch.where(hit, torch.full_like(labels, int(task_cls)), labels)

    @staticmethod
    def _write_back_classes(gt_classes, new_labels, count):
        """The reference relabels the CALLER's class tensors in place (anchors.py:403); so do we: one device->host
        copy for host inputs, one slice copy per tensor otherwise (no per-image synchronisation)."""
        if isinstance(gt_classes, torch.Tensor):
            gt_classes.copy_(new_labels.reshape(gt_classes.shape).to(gt_classes.dtype))
            return
        lens = [int(c.shape[0]) for c in gt_classes]
        host = new_labels.cpu() if any(c.device.type == 'cpu' for c in gt_classes) else None
        for i, (c, n) in enumerate(zip(gt_classes, lens)):
            if n:
                src = host[i, :n] if c.device.type == 'cpu' else new_labels[i, :n]
                c.copy_(src.reshape(c.shape).to(c.dtype))

    def _recycle(self, ws):
        """Called by the fused loss for a transient batch: its kernel has zeroed the keys it read."""
        self._clean_ws.setdefault((ws.device, ws.numel()), []).append(ws)

    def assign(self, gt_boxes, gt_classes, filter_valid=True, task_cls=None, normalizer_out=None, transient=False) -> LabelBatch:
        """Run the assignment kernels; the result feeds either ``targets()`` or the fused loss.
        ``normalizer_out`` (float32 [1], same device) receives sum(num_positives) + 1 in place.
        ``transient``: the batch will be consumed by exactly one fused loss call (what ``DetBenchTrain`` does); that
        kernel then zeroes the assignment keys it read, and the next ``assign`` reuses the workspace without the
        8-bytes-per-anchor memset."""
        boxes, labels, count = self._pack(gt_boxes, gt_classes, filter_valid)
        if task_cls is not None:
            labels = self._relabel_task_cls(boxes, labels, count, task_cls)
            self._write_back_classes(gt_classes, labels, count)
        dev = boxes.device
        B, M = boxes.shape[0], boxes.shape[1]
        lib = _lib.lib()
        anc, desc, gen = self._anchor_tables(dev)
        if gen is not None and (gen.dtype != torch.float64 or not self.use_anchor_generator):
            gen = None
        A = anc.shape[0]
        apad = lib.odk_planar_stride(A)
        num_pos = torch.empty((B,), dtype=torch.float32, device=dev)
        hw = self.anchors.level_hw()
        na = self.anchors.get_anchors_per_location()
        thr = float(np.float32(self.match_threshold))
        with torch.cuda.device(dev):
            if self.use_grid_kernel and desc is not None and thr > 0.0:
                # the assignment stays in the workspace as 64-bit keys: the fused loss reads them directly,
                # int32 `match` is only materialised if somebody asks for it (LabelBatch.match / targets())
                ws_bytes = lib.odk_assign_grid_workspace_bytes(B, A)
                n64 = (ws_bytes + 15) // 16 * 2
                pool = self._clean_ws.get((dev, n64))
                ws, flags = (pool.pop(), 1) if pool else (torch.empty(n64, dtype=torch.int64, device=dev), 0)
                normalizer = torch.empty((1,), dtype=torch.float32, device=dev) if normalizer_out is None else normalizer_out
                if normalizer.dtype != torch.float32 or normalizer.numel() != 1 or normalizer.device != dev:
                    raise ValueError('normalizer_out must be one float32 element on the gt device')
                _lib.check(lib.odk_assign_grid(_lib.ptr(anc), _lib.ptr(desc), _lib.ptr(gen), desc.shape[0], _lib.ptr(boxes),
                                               _lib.ptr(labels), _lib.ptr(count), B, M, _lib.int_array(hw), len(hw), na,
                                               thr, int(bool(filter_valid)), None, _lib.ptr(num_pos),
                                               _lib.ptr(normalizer), flags, _lib.ptr(ws), ws.numel() * 8,
                                               _lib.stream_ptr(dev)))
                keys = ws[:B * apad].view(B, apad)
                return LabelBatch(self, boxes, labels, None, num_pos, keys=keys, normalizer=normalizer, workspace=ws,
                                  transient=bool(transient))
            else:
                match = torch.empty((B, apad), dtype=torch.int32, device=dev)
                ws_bytes = lib.odk_assign_workspace_bytes(B, M)
                ws = torch.empty((ws_bytes + 15) // 16 * 2, dtype=torch.int64, device=dev)
                _lib.check(lib.odk_assign(_lib.ptr(anc), _lib.ptr(boxes), _lib.ptr(labels), _lib.ptr(count), B, M,
                                          _lib.int_array(hw), len(hw), na, thr, int(bool(filter_valid)),
                                          _lib.ptr(match), _lib.ptr(num_pos), _lib.ptr(ws), ws.numel() * 8,
                                          _lib.stream_ptr(dev)))
        if normalizer_out is not None:   # this kernel has no normaliser output: loss.py:261 in torch
            normalizer_out.copy_((num_pos.sum() + 1.0).reshape(normalizer_out.shape))
        return LabelBatch(self, boxes, labels, match, num_pos, normalizer=normalizer_out)

    def _materialize(self, lb: LabelBatch):
        lib = _lib.lib()
        dev = lb.match.device
        B, M = lb.gt_boxes.shape[0], lb.gt_boxes.shape[1]
        A = self.anchors.boxes.shape[0]
        na = self.anchors.get_anchors_per_location()
        cls_flat = torch.empty((B * A,), dtype=torch.int64, device=dev)
        box_flat = torch.empty((B * A * 4,), dtype=torch.float32, device=dev)
        hw = self.anchors.level_hw()
        anc = self._anchor_tables(dev)[0]
        with torch.cuda.device(dev):
            _lib.check(lib.odk_targets(_lib.ptr(anc), _lib.ptr(lb.gt_boxes), _lib.ptr(lb.gt_labels), B, M,
                                       _lib.int_array(hw), len(hw), na, _lib.ptr(lb.match), _lib.ptr(cls_flat),
                                       _lib.ptr(box_flat), _lib.stream_ptr(dev)))
        if self.host_results:   # anchors live on the host: the caller gets host tensors back, like from the reference
            cls_flat, box_flat = cls_flat.cpu(), box_flat.cpu()
        cls_out, box_out = [], []
        offs = self.anchors.level_offsets()
        for i, level in enumerate(range(self.anchors.min_level, self.anchors.max_level + 1)):
            h, w = self.anchors.feat_sizes[level]
            lo, hi = B * offs[i], B * offs[i + 1]
            cls_out.append(cls_flat[lo:hi].view(B, h, w, na))
            box_out.append(box_flat[lo * 4:hi * 4].view(B, h, w, na * 4))
        return cls_out, box_out

    # ---- reference API -----------------------------------------------------------------------
    def label_anchors(self, gt_boxes, gt_classes, filter_valid=True):
        """One image: ([H_l, W_l, na] int64 per level, [H_l, W_l, na*4] fp32 per level, num_positives)."""
        lb = self.assign([gt_boxes], [gt_classes.reshape(-1)], filter_valid=filter_valid)
        cls_t, box_t = lb.targets()
        npos = lb.num_positives.cpu() if self.host_results else lb.num_positives
        return [c[0] for c in cls_t], [b[0] for b in box_t], npos[0]

    def batch_label_anchors(self, gt_boxes, gt_classes, filter_valid=True, task_cls=None):
        """([B, H_l, W_l, na] int64 per level, [B, H_l, W_l, na*4] fp32 per level, num_positives [B])."""
        assert len(gt_boxes) == len(gt_classes)
        lb = self.assign(gt_boxes, gt_classes, filter_valid=filter_valid, task_cls=task_cls)
        cls_t, box_t = lb.targets()
        return cls_t, box_t, (lb.num_positives.cpu() if self.host_results else lb.num_positives)


# ------------------------------------------------------------------------------- detections
def decode_box_outputs(rel_codes, anchors, output_xyxy: bool = False):
    """Relative box codes -> absolute boxes (reference effdet/anchors.py:51-85).

    Plain elementwise torch on whatever device the inputs live on; the fused decode used by
    the detection path is inside odk_detect."""
    ycenter_a = (anchors[:, 0] + anchors[:, 2]) / 2
    xcenter_a = (anchors[:, 1] + anchors[:, 3]) / 2
    ha = anchors[:, 2] - anchors[:, 0]
    wa = anchors[:, 3] - anchors[:, 1]
    ty, tx, th, tw = rel_codes.unbind(dim=1)
    w = torch.exp(tw) * wa
    h = torch.exp(th) * ha
    yc = ty * ha + ycenter_a
    xc = tx * wa + xcenter_a
    ymin, xmin, ymax, xmax = yc - h / 2., xc - w / 2., yc + h / 2., xc + w / 2.
    order = [xmin, ymin, xmax, ymax] if output_xyxy else [ymin, xmin, ymax, xmax]
    return torch.stack(order, dim=1)


def clip_boxes_xyxy(boxes: torch.Tensor, size: torch.Tensor):
    """Clamp xyxy boxes to [0, size] (reference effdet/anchors.py:88-92)."""
    return boxes.clamp(min=0).min(torch.cat([size, size], dim=0))


def detect_batch(cls_topk, box_topk, anchor_boxes, indices, classes, img_scale=None, img_size=None,
                 max_det_per_image: int = 100, soft_nms: bool = False):
    """Batched odk_detect: -> dets [B, D, 6] zero padded, count [B] int32, src [B, D] int32."""
    lib = _lib.lib()
    _lib.require_cuda(cls_topk, 'cls_outputs')
    dev = cls_topk.device
    B, N = indices.shape[0], indices.shape[1]
    cls_topk = cls_topk.reshape(B, N).float().contiguous()
    box_topk = box_topk.reshape(B, N, 4).float().contiguous()
    indices = indices.to(torch.int64).contiguous()
    classes = classes.to(torch.int64).contiguous()
    anchor_boxes = anchor_boxes.to(dev, torch.float32).contiguous()
    D = int(max_det_per_image)
    dets = torch.empty((B, D, 6), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    src = torch.empty((B, D), dtype=torch.int32, device=dev)
    scale = None if img_scale is None else img_scale.to(dev, torch.float32).reshape(B).contiguous()
    size = None if img_size is None else img_size.to(dev, torch.float32).reshape(B, 2).contiguous()
    params = _lib.DetectParams(D, int(bool(soft_nms)), float(np.float32(0.01)), 0.3, 0.5, 0.3, float(np.float32(0.001)))
    with torch.cuda.device(dev):
        _lib.check(lib.odk_detect(_lib.ptr(cls_topk), _lib.ptr(box_topk), _lib.ptr(indices), _lib.ptr(classes), B, N,
                                  _lib.ptr(anchor_boxes), anchor_boxes.shape[0], _lib.ptr(scale), _lib.ptr(size),
                                  params, _lib.ptr(dets), _lib.ptr(count), _lib.ptr(src), _lib.stream_ptr(dev)))
    return dets, count, src


def generate_detections(cls_outputs, box_outputs, anchor_boxes, indices, classes,
                        img_scale: Optional[torch.Tensor], img_size: Optional[torch.Tensor],
                        max_det_per_image: int = 100, soft_nms: bool = False):
    """One image's detections [n <= max_det, 6] = x0, y0, x1, y1, score, class (1-based).

    Same contract as the reference (effdet/anchors.py:95-172): inputs are the per-image slices of
    ``_post_process``; rows are NOT padded.  The row count is data dependent, so this call reads
    one int back from the device (the reference syncs several times per image)."""
    assert box_outputs.shape[-1] == 4
    assert anchor_boxes.shape[-1] == 4
    assert cls_outputs.shape[-1] == 1
    scale = None if img_scale is None else img_scale.reshape(1)
    size = None if img_size is None else img_size.reshape(1, 2)
    dets, count, _ = detect_batch(cls_outputs.reshape(1, -1), box_outputs.reshape(1, -1, 4), anchor_boxes,
                                  indices.reshape(1, -1), classes.reshape(1, -1), scale, size,
                                  max_det_per_image, soft_nms)
    return dets[0, :int(count.item())]

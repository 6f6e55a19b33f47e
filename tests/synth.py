"""Seeded synthetic inputs for the dense per-anchor hot path.

Everything here is numpy ``RandomState`` based: those streams are frozen across numpy
versions, so the golden generator (run once, next to the reference) and the tests (run
anywhere) see bit-identical inputs without shipping the inputs themselves.
"""
import numpy as np

ASPECTS = [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)]
NUM_SCALES = 3
MIN_LEVEL, MAX_LEVEL = 3, 7
# name -> (image size, anchor_scale); model_config.py:33-41, :498 in the reference
MODEL_SHAPES = {'d0': (512, 4.0), 'd3': (896, 4.0), 'd5': (1280, 4.0), 'd7': (1536, 5.0)}


def feat_hw(image_size, min_level=MIN_LEVEL, max_level=MAX_LEVEL):
    """[(H_l, W_l)] for levels min..max (same recurrence as anchors.py:175-188)."""
    h = w = image_size
    out = []
    for lvl in range(1, max_level + 1):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        if lvl >= min_level:
            out.append((h, w))
    return out


def num_anchors(image_size):
    return 9 * sum(h * w for h, w in feat_hw(image_size))


def head_outputs(seed, batch, image_size, num_classes, mean=-4.6, std=1.5, box_std=0.2,
                 tie_free=True):
    """Per-level NCHW class logits [B, 9C, H, W] and box regressions [B, 36, H, W], fp32.

    ``tie_free`` makes the top 20000 logits of every image distinct (see ``make_tie_free``),
    because ``torch.topk`` tie order is unspecified.
    """
    rs = np.random.RandomState(seed)
    cls, box = [], []
    for (h, w) in feat_hw(image_size):
        c = (rs.standard_normal((batch, 9 * num_classes, h, w)) * std + mean).astype(np.float32)
        b = (rs.standard_normal((batch, 36, h, w)) * box_std).astype(np.float32)
        cls.append(c)
        box.append(b)
    if tie_free:
        make_tie_free(cls)
    return cls, box


def make_tie_free(cls, ntop=20000):
    """Make the ``ntop`` largest logits of each image distinct, in place.

    ``torch.topk`` returns tied values in an unspecified order, so bit-exact index parity is only
    defined on inputs whose top region is tie free.  Sort the image's values, push each duplicate
    in the top region up to the next float above its predecessor (repeat until strictly
    increasing), scatter back.  Moves a value by a few ulps at most.
    """
    batch = cls[0].shape[0]
    for i in range(batch):
        flat = np.concatenate([c[i].ravel() for c in cls])
        order = np.argsort(flat, kind='stable')
        lo = max(0, flat.size - ntop - 1)
        s = flat[order[lo:]]
        for _ in range(10000):
            dup = np.nonzero(s[1:] <= s[:-1])[0]
            if dup.size == 0:
                break
            s[dup + 1] = np.nextafter(s[dup], np.float32(np.inf))
        assert np.all(s[1:] > s[:-1])
        flat[order[lo:]] = s
        off = 0
        for c in cls:
            n = c[i].size
            c[i] = flat[off:off + n].reshape(c[i].shape)
            off += n


def planted_outputs(seed, batch, image_size, num_classes, n_obj=50, n_per=20):
    """'Sparse' score regime: background mean -7, a few planted object clusters at +3."""
    cls, box = head_outputs(seed, batch, image_size, num_classes, mean=-7.0, std=1.0, tie_free=False)
    rs = np.random.RandomState(seed + 7919)
    sizes = feat_hw(image_size)
    for i in range(batch):
        for _ in range(n_obj):
            lvl = rs.randint(0, 3)
            h, w = sizes[lvl]
            y0, x0 = rs.randint(0, h), rs.randint(0, w)
            c = rs.randint(0, num_classes)
            for _ in range(n_per):
                a = rs.randint(0, 9)
                y = min(h - 1, max(0, y0 + rs.randint(-1, 2)))
                x = min(w - 1, max(0, x0 + rs.randint(-1, 2)))
                cls[lvl][i, a * num_classes + c, y, x] = np.float32(3.0 + rs.standard_normal() * 0.5)
    make_tie_free(cls)
    return cls, box


def gt_boxes(seed, batch, image_size, m, num_classes, integer=False):
    """[B, m, 4] yxyx fp32 boxes and [B, m] int64 1-based classes (SURVEY 8d recipe)."""
    rs = np.random.RandomState(seed)
    s = float(image_size)
    cy, cx = rs.uniform(0, s, (batch, m)), rs.uniform(0, s, (batch, m))
    hh = rs.uniform(8, 0.4 * s + 8, (batch, m))
    ww = rs.uniform(8, 0.4 * s + 8, (batch, m))
    b = np.stack([cy - hh / 2, cx - ww / 2, cy + hh / 2, cx + ww / 2], -1)
    b = np.clip(b, 0, s)
    if integer:
        b = np.round(b / 8) * 8
    cls = rs.randint(1, num_classes + 1, (batch, m)).astype(np.int64)
    return b.astype(np.float32), cls


def nms_candidates(seed, n, image_size, num_classes, n_clusters=40):
    """Clustered xyxy boxes / scores / classes that give NMS real work to do."""
    rs = np.random.RandomState(seed)
    s = float(image_size)
    cc = rs.uniform(0.1 * s, 0.9 * s, (n_clusters, 2))
    wh = rs.uniform(0.05 * s, 0.3 * s, (n_clusters, 2))
    kcls = rs.randint(0, num_classes, n_clusters)
    which = rs.randint(0, n_clusters, n)
    ctr = cc[which] + rs.standard_normal((n, 2)) * 0.03 * s
    sz = wh[which] * np.exp(rs.standard_normal((n, 2)) * 0.15)
    boxes = np.concatenate([ctr - sz / 2, ctr + sz / 2], 1).astype(np.float32)
    scores = rs.uniform(0.02, 1.0, n).astype(np.float32)
    classes = np.where(rs.uniform(size=n) < 0.8, kcls[which], rs.randint(0, num_classes, n)).astype(np.int64)
    return boxes, scores, classes


# (tag, seed, detections, gt boxes, classes, nms_iou, nms_max, difficult / group-of flags): per-image evaluation cases
EVAL_CASES = [('plain', 71, 100, 12, 8, 1.0, 10000, False), ('flags', 72, 100, 30, 5, 1.0, 10000, True),
              ('nms', 73, 100, 10, 4, 0.3, 50, True), ('one_class', 74, 60, 6, 1, 1.0, 10000, False),
              ('no_gt_of_class', 75, 40, 3, 20, 1.0, 10000, False)]


def eval_case(seed, n_det, n_gt, num_classes, size=512.0):
    """Detections scattered around (and away from) the gt boxes, unique scores, a few invalid boxes."""
    rs = np.random.RandomState(seed)
    gb, gc = gt_boxes(seed, 1, int(size), n_gt, num_classes)
    gtb, gt_classes = gb[0], (gc[0] - 1).astype(np.int64)               # 0-based like the evaluator sees them
    src = rs.randint(0, max(n_gt, 1), n_det)
    det = gtb[src] + rs.standard_normal((n_det, 4)).astype(np.float32) * rs.choice([2.0, 12.0, 60.0], (n_det, 1)).astype(np.float32)
    det = det.astype(np.float32)
    det[::17, 2] = det[::17, 0]                                               # invalid: ymax == ymin
    cls = np.where(rs.uniform(size=n_det) < 0.8, gt_classes[src], rs.randint(0, num_classes, n_det)).astype(np.int64)
    scores = rs.permutation(n_det).astype(np.float32) / np.float32(n_det) * np.float32(0.98) + np.float32(0.01)
    difficult = rs.uniform(size=n_gt) < 0.15
    group_of = rs.uniform(size=n_gt) < 0.15
    return det, scores, cls, gtb, gt_classes, difficult, group_of



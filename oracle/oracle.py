"""Python face of the CPU oracle (ctypes over oracle/liboracle.so + numpy glue).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; never from ood_object_detection_b200/.  Parity is pinned
against tests/golden/*.npz (outputs of the unmodified reference python, see
tests/golden/make_golden.py) by tests/test_oracle_golden.py.

Function names follow the reference API they restate (file:line cited per function).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build():
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    subprocess.run(['make', '-C', _HERE, 'liboracle.so'], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_soft_nms.restype = ctypes.c_int64
        _LIB.orc_generate_detections.restype = ctypes.c_int64
        _LIB.orc_nms.restype = ctypes.c_int64
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """OpenMP threads = cores this process may run on (launchers such as torchrun export
    OMP_NUM_THREADS=1, which would turn the CPU baseline into a single-thread run)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(ctypes.c_int(n))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------ anchors
def get_feat_sizes(image_size, max_level):
    """effdet/anchors.py:175-188."""
    fs = [tuple(image_size)]
    for _ in range(max_level):
        h, w = fs[-1]
        fs.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return fs


def anchor_boxes(min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size):
    """effdet/anchors.py:249-299: float64 grid + half sizes, cast to fp32; rows ordered
    level -> y -> x -> (octave, aspect)."""
    fs = get_feat_sizes(image_size, max_level)
    if not isinstance(anchor_scale, (list, tuple)):
        anchor_scale = [anchor_scale] * (max_level - min_level + 1)
    out = []
    for level in range(min_level, max_level + 1):
        stride = (fs[0][0] // fs[level][0], fs[0][1] // fs[level][1])
        per_cfg = []
        for octave in range(num_scales):
            for aspect in aspect_ratios:
                sc = anchor_scale[level - min_level]
                bx = sc * stride[1] * 2 ** (octave / float(num_scales))
                by = sc * stride[0] * 2 ** (octave / float(num_scales))
                if isinstance(aspect, (list, tuple)):
                    ax, ay = aspect
                else:
                    ax = np.sqrt(aspect)
                    ay = 1.0 / ax
                hx, hy = bx * ax / 2.0, by * ay / 2.0
                x = np.arange(stride[1] / 2, image_size[1], stride[1])
                y = np.arange(stride[0] / 2, image_size[0], stride[0])
                xv, yv = np.meshgrid(x, y)
                xv, yv = xv.reshape(-1), yv.reshape(-1)
                per_cfg.append(np.stack([yv - hy, xv - hx, yv + hy, xv + hx], 1)[:, None, :])
        out.append(np.concatenate(per_cfg, 1).reshape(-1, 4))
    return np.vstack(out).astype(np.float32)


# ------------------------------------------------------------------------------------ labeler
def iou_matrix(b1, b2):
    """IouSimilarity.compare (region_similarity_calculator.py:59-101) -> [N, M]."""
    b1, b2 = _f32(b1).reshape(-1, 4), _f32(b2).reshape(-1, 4)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    lib().orc_iou_matrix(_p(b1, _f32p), ctypes.c_int64(b1.shape[0]), _p(b2, _f32p), ctypes.c_int64(b2.shape[0]),
                         _p(out, _f32p))
    return out


def batch_label_anchors(anchors, gt_boxes, gt_classes, match_threshold=0.5, filter_valid=True, task_cls=None):
    """AnchorLabeler.batch_label_anchors (effdet/anchors.py:384-438) on flat [B, A] outputs.

    gt_boxes / gt_classes: sequences (or arrays) of per-image [M_i,4] / [M_i].  Returns
    cls_targets [B, A] int64, box_targets [B, A, 4] fp32, num_positives [B] fp32, match [B, A] and
    the (possibly task_cls-relabelled) classes list.  Per-level views are a reshape away:
    level l covers rows [9*sum(HW_<l), 9*sum(HW_<=l)) in (y, x, anchor) order.
    """
    anchors = _f32(anchors)
    A = anchors.shape[0]
    B = len(gt_boxes)
    boxes_f, labels_f = [], []
    classes_out = []
    for i in range(B):
        gb = _f32(np.asarray(gt_boxes[i])).reshape(-1, 4)
        gc = np.array(gt_classes[i]).reshape(-1).copy()
        if task_cls is not None:  # anchors.py:396-403
            tmask = gc == task_cls
            if (~tmask).sum() > 0:
                sims = iou_matrix(gb[tmask], gb)
                if sims.shape[0] > 0:
                    overl = (sims > np.float32(0.9)).max(0)
                else:
                    overl = np.zeros(gb.shape[0], bool)
                gc[overl] = task_cls
        classes_out.append(gc)
        if filter_valid:  # anchors.py:405-408
            v = gc > -1
            gb, gc = gb[v], gc[v]
        boxes_f.append(gb)
        labels_f.append(np.trunc(gc).astype(np.int64) if gc.dtype.kind == 'f' else gc.astype(np.int64))
    Mmax = max([1] + [b.shape[0] for b in boxes_f])
    gtb = np.zeros((B, Mmax, 4), np.float32)
    gtl = np.zeros((B, Mmax), np.int64)
    cnt = np.zeros((B,), np.int64)
    for i in range(B):
        m = boxes_f[i].shape[0]
        gtb[i, :m], gtl[i, :m], cnt[i] = boxes_f[i], labels_f[i], m
    match = np.empty((B, A), np.int64)
    cls_t = np.empty((B, A), np.int64)
    box_t = np.empty((B, A, 4), np.float32)
    npos = np.empty((B,), np.float32)
    thr = np.float32(match_threshold)
    lib().orc_assign_batch(_p(anchors, _f32p), ctypes.c_int64(A), _p(gtb, _f32p), _p(gtl, _i64p), _p(cnt, _i64p),
                           ctypes.c_int64(B), ctypes.c_int64(Mmax), ctypes.c_float(thr), ctypes.c_float(thr),
                           ctypes.c_int(1), ctypes.c_int(1), _p(match, _i64p), _p(cls_t, _i64p), _p(box_t, _f32p),
                           _p(npos, _f32p))
    return cls_t, box_t, npos, match, classes_out


def split_levels(flat, feat_hw, na=9):
    """[B, A, ...] -> per-level [B, H, W, na*(...)] like anchors.py:420-432."""
    out, off = [], 0
    B = flat.shape[0]
    for (h, w) in feat_hw:
        n = h * w * na
        out.append(np.ascontiguousarray(flat[:, off:off + n]).reshape(B, h, w, -1))
        off += n
    return out


# ------------------------------------------------------------------------------------ loss
def loss_fn(cls_outputs, box_outputs, cls_targets, box_targets, num_positives, num_classes, alpha, gamma, delta,
            box_loss_weight, label_smoothing=0.0, legacy_focal=False, want_grad=False):
    """effdet/loss.py:224-298.  Per-level lists: cls_outputs [B,9C,H,W], box_outputs [B,36,H,W],
    cls_targets [B,H,W,9] int, box_targets [B,H,W,36].  Returns (total, cls, box) as python
    floats (double accumulation) and, if want_grad, d total / d outputs per level."""
    nps = np.float32(np.float32(np.sum(_f32(num_positives), dtype=np.float32)) + np.float32(1.0))
    cls_sum = box_sum = 0.0
    gcs, gbs = [], []
    for l in range(len(cls_outputs)):
        co, bo = _f32(cls_outputs[l]), _f32(box_outputs[l])
        ct, bt = _i64(cls_targets[l]), _f32(box_targets[l])
        B, _, H, W = co.shape
        na = bo.shape[1] // 4
        gc = np.empty_like(co) if want_grad else None
        gb = np.empty_like(bo) if want_grad else None
        cs, bs = ctypes.c_double(0), ctypes.c_double(0)
        lib().orc_loss_level(_p(co, _f32p), _p(bo, _f32p), _p(ct, _i64p), _p(bt, _f32p), ctypes.c_int64(B),
                             ctypes.c_int64(H), ctypes.c_int64(W), ctypes.c_int64(num_classes), ctypes.c_int64(na),
                             ctypes.c_float(nps), ctypes.c_float(alpha), ctypes.c_float(gamma), ctypes.c_float(delta),
                             ctypes.c_float(label_smoothing), ctypes.c_int(int(legacy_focal)),
                             ctypes.c_float(box_loss_weight), ctypes.byref(cs), ctypes.byref(bs),
                             _p(gc, _f32p) if want_grad else None, _p(gb, _f32p) if want_grad else None)
        cls_sum += cs.value
        box_sum += bs.value
        gcs.append(gc)
        gbs.append(gb)
    total = cls_sum + box_loss_weight * box_sum
    if want_grad:
        return total, cls_sum, box_sum, gcs, gbs
    return total, cls_sum, box_sum


# ------------------------------------------------------------------------------------ post-process
def post_process(cls_outputs, box_outputs, num_levels, num_classes, max_detection_points=5000):
    """effdet/bench.py:12-56 -> (cls [B,K,1], box [B,K,4], indices [B,K], classes [B,K])."""
    co = [_f32(c) for c in cls_outputs[:num_levels]]
    bo = [_f32(b) for b in box_outputs[:num_levels]]
    B = co[0].shape[0]
    na = bo[0].shape[1] // 4
    hw = np.array([c.shape[2] * c.shape[3] for c in co], np.int64)
    K = int(max_detection_points)
    cp = (_f32p * num_levels)(*[_p(c, _f32p) for c in co])
    bp = (_f32p * num_levels)(*[_p(b, _f32p) for b in bo])
    cls_k = np.empty((B, K), np.float32)
    box_k = np.empty((B, K, 4), np.float32)
    idx = np.empty((B, K), np.int64)
    klass = np.empty((B, K), np.int64)
    lib().orc_post_process(cp, bp, _p(hw, _i64p), ctypes.c_int(num_levels), ctypes.c_int64(B),
                           ctypes.c_int64(num_classes), ctypes.c_int64(na), ctypes.c_int64(K), _p(cls_k, _f32p),
                           _p(box_k, _f32p), _p(idx, _i64p), _p(klass, _i64p))
    return cls_k[:, :, None], box_k, idx, klass


def generate_detections(cls_outputs, box_outputs, anchor_boxes_, indices, classes, img_scale=None, img_size=None,
                        max_det_per_image=100, soft_nms=False, return_src=False):
    """effdet/anchors.py:95-172 for one image -> [n<=D, 6] (x0,y0,x1,y1,score,class)."""
    cls = _f32(cls_outputs).reshape(-1)
    box = _f32(box_outputs).reshape(-1, 4)
    anc = _f32(anchor_boxes_)
    ind, kc = _i64(indices), _i64(classes)
    N, D = cls.shape[0], int(max_det_per_image)
    det = np.zeros((D, 6), np.float32)
    src = np.zeros((D,), np.int64)
    size = _f32(img_size if img_size is not None else [0, 0])
    n = lib().orc_generate_detections(
        _p(cls, _f32p), _p(box, _f32p), _p(anc, _f32p), _p(ind, _i64p), _p(kc, _i64p), ctypes.c_int64(N),
        ctypes.c_int(img_scale is not None), ctypes.c_float(0.0 if img_scale is None else float(img_scale)),
        ctypes.c_int(img_size is not None), _p(size, _f32p), ctypes.c_int64(D), ctypes.c_int(int(soft_nms)),
        ctypes.c_float(np.float32(0.01)), ctypes.c_double(0.3), ctypes.c_float(0.5), ctypes.c_float(0.3),
        ctypes.c_float(np.float32(0.001)), _p(det, _f32p), _p(src, _i64p))
    if return_src:
        return det[:n], src[:n]
    return det[:n]


def soft_nms(boxes, scores, method_gaussian=True, sigma=0.5, iou_threshold=0.5, score_threshold=0.005,
             max_rounds=-1):
    """effdet/soft_nms.py:42-112 -> (kept indices int64, rescored values)."""
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores)
    n = s.shape[0]
    idx = np.zeros((max(n, 1),), np.int64)
    sc = np.zeros((max(n, 1),), np.float32)
    c = lib().orc_soft_nms(_p(b, _f32p), _p(s, _f32p), ctypes.c_int64(n), ctypes.c_int(int(method_gaussian)),
                           ctypes.c_float(sigma), ctypes.c_float(iou_threshold), ctypes.c_float(score_threshold),
                           ctypes.c_int64(max_rounds), _p(idx, _i64p), _p(sc, _f32p))
    return idx[:c], sc[:c]


def batched_soft_nms(boxes, scores, idxs, method_gaussian=True, sigma=0.5, iou_threshold=0.5,
                     score_threshold=0.001, max_rounds=-1):
    """effdet/soft_nms.py:115-169 (class offsets added in fp32, then soft_nms)."""
    b = _f32(boxes).reshape(-1, 4)
    if b.size == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), np.float32)
    off = _i64(idxs).astype(np.float32) * (b.max() + np.float32(1))
    return soft_nms(b + off[:, None], scores, method_gaussian, sigma, iou_threshold, score_threshold, max_rounds)


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms, CPU kernel semantics."""
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores)
    keep = np.zeros((max(s.shape[0], 1),), np.int64)
    n = lib().orc_nms(_p(b, _f32p), _p(s, _f32p), ctypes.c_int64(s.shape[0]), ctypes.c_double(iou_threshold),
                      _p(keep, _i64p))
    return keep[:n]


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.boxes._batched_nms_coordinate_trick (ops/boxes.py:94-111)."""
    b = _f32(boxes).reshape(-1, 4)
    if b.size == 0:
        return np.zeros((0,), np.int64)
    off = _i64(idxs).astype(np.float32) * (b.max() + np.float32(1))
    return nms(b + off[:, None], scores, iou_threshold)


def decode_box_outputs(rel_codes, anchors, output_xyxy=False):
    """effdet/anchors.py:51-85."""
    r, a = _f32(rel_codes).reshape(-1, 4), _f32(anchors).reshape(-1, 4)
    out = np.empty_like(r)
    lib().orc_decode(_p(r, _f32p), _p(a, _f32p), ctypes.c_int64(r.shape[0]), ctypes.c_int(int(output_xyxy)),
                     _p(out, _f32p))
    return out


def ood_scores(rows, temperature=1.0):
    """energy = -T*logsumexp(row/T), max_logit = max(row) (not in the reference)."""
    r = _f32(rows)
    r2 = r.reshape(-1, r.shape[-1])
    e = np.empty((r2.shape[0],), np.float32)
    m = np.empty((r2.shape[0],), np.float32)
    lib().orc_ood(_p(r2, _f32p), ctypes.c_int64(r2.shape[0]), ctypes.c_int64(r2.shape[1]),
                  ctypes.c_float(temperature), _p(e, _f32p), _p(m, _f32p))
    return e.reshape(r.shape[:-1]), m.reshape(r.shape[:-1])


def gather_logit_rows(cls_outputs, anchor_idx, num_classes, na=9):
    """The [.., C] rows the reference gathers at bench.py:51-52, straight from the NCHW levels."""
    B = cls_outputs[0].shape[0]
    allc = np.concatenate([np.transpose(c, (0, 2, 3, 1)).reshape(B, -1, num_classes) for c in cls_outputs], 1)
    return np.take_along_axis(allc, np.asarray(anchor_idx)[:, :, None].astype(np.int64), 1)


# ------------------------------------------------------------------------------------ evaluation
def match_detections(det_boxes, det_scores, det_classes, gt_boxes, gt_classes, num_classes, gt_difficult=None,
                     gt_group_of=None, match_iou=0.5, nms_iou=1.0, nms_max=10000):
    """effdet/evaluation/per_image_evaluation.py:29-92 for one image (boxes yxyx, classes 0-based) ->
    (label [N] int8: 1 tp / 0 fp / -1 ignored / -2 removed, corloc [num_classes] uint8)."""
    db, ds, dc = _f32(det_boxes).reshape(-1, 4), _f32(det_scores).reshape(-1), _i64(det_classes).reshape(-1)
    gb, gc = _f32(gt_boxes).reshape(-1, 4), _i64(gt_classes).reshape(-1)
    N, M = db.shape[0], gb.shape[0]
    u8p = ctypes.POINTER(ctypes.c_uint8)
    dif = None if gt_difficult is None else np.ascontiguousarray(gt_difficult, np.uint8)
    gof = None if gt_group_of is None else np.ascontiguousarray(gt_group_of, np.uint8)
    label = np.empty((max(N, 1),), np.int8)
    corloc = np.zeros((num_classes,), np.uint8)
    lib().orc_match_detections(_p(db, _f32p), _p(ds, _f32p), _p(dc, _i64p), ctypes.c_int64(N), _p(gb, _f32p), _p(gc, _i64p),
                               None if dif is None else _p(dif, u8p), None if gof is None else _p(gof, u8p), ctypes.c_int64(M),
                               ctypes.c_int64(num_classes), ctypes.c_double(match_iou), ctypes.c_double(nms_iou),
                               ctypes.c_int64(nms_max), label.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), _p(corloc, u8p))
    return label[:N], corloc

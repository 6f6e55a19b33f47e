"""Generate golden input/output vectors by running the UNMODIFIED reference python path.

Run once in the build container (the reference tree cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``effdet.*`` straight from /root/reference (a namespace package, so only the hot-path
modules get imported) with the two out-of-tree shims SURVEY.md section 8c documents:
  * ``TargetAssigner._create_regression_targets`` result made contiguous (torch>=2 ``.view``),
  * ``effdet.anchors.batched_nms`` pinned to torchvision's coordinate-trick path (the path the
    reference takes on CUDA for n<=5000).
Outputs land in tests/golden/*.npz.  Inputs come from tests/synth.py seeds and are not stored
unless they are tiny hand-written known-answer cases.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

import torch  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

from effdet.object_detection.target_assigner import TargetAssigner  # noqa: E402

_orig = TargetAssigner._create_regression_targets
TargetAssigner._create_regression_targets = lambda self, *a, **k: _orig(self, *a, **k).contiguous()
import effdet.anchors as RA  # noqa: E402
import torchvision.ops.boxes as tvb  # noqa: E402
from effdet.bench import _post_process  # noqa: E402  (import first: jit.script compiles at import)

RA.batched_nms = tvb._batched_nms_coordinate_trick
from effdet.anchors import Anchors, AnchorLabeler, generate_detections, decode_box_outputs  # noqa: E402
from effdet.loss import loss_fn  # noqa: E402
from effdet.soft_nms import soft_nms, batched_soft_nms  # noqa: E402

import synth  # noqa: E402


def mk_anchors(size, scale=4.0):
    return Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size))


def save(name, **kw):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in kw.items()})
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


# ----------------------------------------------------------------------------- anchors
def gold_anchors():
    out = {}
    for name, (size, scale) in synth.MODEL_SHAPES.items():
        b = mk_anchors(size, scale).boxes.numpy()
        out[f'{name}_count'] = b.shape[0]
        out[f'{name}_head'] = b[:32]
        out[f'{name}_tail'] = b[-32:]
        out[f'{name}_sum64'] = b.astype(np.float64).sum(0)
        out[f'{name}_abs64'] = np.abs(b.astype(np.float64)).sum()
        # position-weighted checksum so a permutation of rows is caught
        wgt = (np.arange(b.shape[0], dtype=np.float64) % 1009 + 1)[:, None]
        out[f'{name}_wsum64'] = (b.astype(np.float64) * wgt).sum(0)
    out['s128'] = mk_anchors(128).boxes.numpy()
    out['s256_scale3'] = Anchors(3, 7, 3, synth.ASPECTS, 3.0, (256, 256)).boxes.numpy()
    save('anchors', **out)


# ----------------------------------------------------------------------------- labeler
def run_label(labeler, boxes, classes, **kw):
    cls_t, box_t, npos = labeler.batch_label_anchors(boxes, classes, **kw)
    cls_flat = torch.cat([c.reshape(c.shape[0], -1) for c in cls_t], 1)
    box_flat = torch.cat([b.reshape(b.shape[0], -1, 4) for b in box_t], 1)
    return cls_flat, box_flat, npos


def gold_labeler():
    out = {}
    a512 = mk_anchors(512)
    lab = AnchorLabeler(a512, 90, match_threshold=0.5)
    f = torch.float32

    def kat(tag, boxes, classes, **kw):
        c, b, n = run_label(lab, boxes, classes, **kw)
        pos = torch.nonzero(c[0] != -1).flatten()
        out[f'kat_{tag}_boxes'] = boxes[0] if isinstance(boxes, list) else boxes[0]
        out[f'kat_{tag}_classes'] = classes[0]
        out[f'kat_{tag}_pos_idx'] = pos
        out[f'kat_{tag}_pos_cls'] = c[0][pos]
        out[f'kat_{tag}_pos_box'] = b[0][pos]
        out[f'kat_{tag}_npos'] = n
        # everything that is not positive must be exactly (-1, 0)
        neg = torch.ones(c.shape[1], dtype=torch.bool)
        neg[pos] = False
        assert (c[0][neg] == -1).all() and (b[0][neg] == 0).all()

    kat('empty', [torch.zeros((0, 4), dtype=f)], [torch.zeros((0,), dtype=torch.int64)])
    kat('zero_iou', [torch.tensor([[5000., 5000, 5100, 5100], [100, 100, 100, 100]], dtype=f)],
        [torch.tensor([7, 9])])
    kat('identical', [torch.tensor([[100., 100, 200, 200], [100, 100, 200, 200]], dtype=f)],
        [torch.tensor([3, 5])])
    kat('tiny', [torch.tensor([[250., 250, 253, 253]], dtype=f)], [torch.tensor([4])])
    pb = -torch.ones((1, 100, 4), dtype=f)
    pc = -torch.ones((1, 100), dtype=f)
    pb[0, 0] = torch.tensor([100., 100, 200, 200])
    pb[0, 1] = torch.tensor([100., 100, 200, 200])
    pc[0, 0], pc[0, 1] = 3, 5
    kat('padded_float', pb, pc)
    # two gts whose best anchor is the same one (forced-match collision: lowest gt wins) plus
    # grid-aligned gts that tie across many anchors
    kat('collide', [torch.tensor([[250., 250, 253, 253], [250.5, 250.5, 253.5, 253.5],
                                   [64, 64, 192, 192], [0, 0, 512, 512]], dtype=f)],
        [torch.tensor([4, 8, 2, 1])])
    # filter_valid=False with a -1 label -> class target -2 ("ignore") appears
    kat('nofilter', [torch.tensor([[100., 100, 200, 200], [300, 300, 420, 400]], dtype=f)],
        [torch.tensor([-1, 6])], filter_valid=False)

    for tag, size, m, seed, integer in [('r256_m10', 256, 10, 11, False), ('r256_m100', 256, 100, 12, False),
                                        ('r256_int', 256, 24, 13, True), ('r512_m10', 512, 10, 14, False)]:
        anc = mk_anchors(size)
        lb = AnchorLabeler(anc, 90, match_threshold=0.5)
        gb, gc = synth.gt_boxes(seed, 3, size, m, 90, integer=integer)
        gb, gc = torch.from_numpy(gb), torch.from_numpy(gc)
        gc[1, m // 2:] = -1  # ragged validity
        if tag == 'r256_m10':
            gc[2, :] = -1  # an image with no valid gt
        c, b, n = run_label(lb, gb, gc)
        out[f'{tag}_cls'] = c.to(torch.int16)
        out[f'{tag}_box'] = b
        out[f'{tag}_npos'] = n
        out[f'{tag}_gc'] = gc

    # other match thresholds (compared in fp32 by torch)
    anc = mk_anchors(256)
    gb, gc = synth.gt_boxes(15, 2, 256, 20, 90)
    for thr in (0.4, 0.7):
        lb = AnchorLabeler(anc, 90, match_threshold=thr)
        c, b, n = run_label(lb, torch.from_numpy(gb), torch.from_numpy(gc))
        out[f'thr{int(thr * 10)}_cls'] = c.to(torch.int16)
        out[f'thr{int(thr * 10)}_box'] = b
        out[f'thr{int(thr * 10)}_npos'] = n

    # task_cls relabel (anchors.py:396-403): mutates gt classes in place
    gb = torch.tensor([[[50., 50, 150, 150], [52, 51, 150, 151], [10, 10, 60, 60], [49, 50, 151, 150]]], dtype=f)
    gc = torch.tensor([[5, 9, 9, 7]])
    gcm = gc.clone()
    lb = AnchorLabeler(anc, 90, match_threshold=0.5)
    c, b, n = run_label(lb, gb, gcm, task_cls=5)
    out['task_boxes'], out['task_classes_in'], out['task_classes_out'] = gb, gc, gcm
    out['task_cls'] = c.to(torch.int16)
    out['task_box'] = b
    out['task_npos'] = n
    save('labeler', **out)


# ----------------------------------------------------------------------------- loss
def gold_loss():
    out = {}
    cases = [
        # tag, size, B, C, M, alpha, gamma, delta, w, smoothing, legacy, plant_ignore
        ('new_c90', 128, 3, 90, 6, 0.25, 1.5, 0.1, 50.0, 0.0, False, False),
        ('new_c1', 128, 2, 1, 6, 0.15, 0.0, 0.1, 5.0, 0.0, False, False),
        ('new_smooth', 128, 2, 7, 6, 0.25, 1.5, 0.1, 50.0, 0.1, False, True),
        ('legacy', 128, 2, 20, 6, 0.25, 1.5, 0.1, 50.0, 0.0, True, True),
        ('legacy_g0', 128, 2, 5, 6, 0.25, 0.0, 0.2, 1.0, 0.0, True, False),
        ('new_256', 256, 2, 90, 12, 0.25, 1.5, 0.1, 50.0, 0.0, False, False),
    ]
    for i, (tag, size, B, C, M, alpha, gamma, delta, w, sm, legacy, plant) in enumerate(cases):
        anc = mk_anchors(size)
        lb = AnchorLabeler(anc, C, match_threshold=0.5)
        gb, gc = synth.gt_boxes(100 + i, B, size, M, C)
        cls_t, box_t, npos = lb.batch_label_anchors(torch.from_numpy(gb), torch.from_numpy(gc))
        if plant:
            rs = np.random.RandomState(500 + i)
            for t in cls_t:
                m = torch.from_numpy(rs.uniform(size=tuple(t.shape)) < 0.02)
                t[m] = -2
        co, bo = synth.head_outputs(200 + i, B, size, C, tie_free=False)
        co = [torch.from_numpy(x).requires_grad_(True) for x in co]
        bo = [torch.from_numpy(x).requires_grad_(True) for x in bo]
        tot, cl, bl = loss_fn(co, bo, cls_t, box_t, npos, num_classes=C, alpha=alpha, gamma=gamma, delta=delta,
                              box_loss_weight=w, label_smoothing=sm, legacy_focal=legacy)
        tot.backward()
        out[f'{tag}_params'] = np.array([size, B, C, M, alpha, gamma, delta, w, sm, float(legacy), float(plant),
                                         100 + i, 200 + i, 500 + i], dtype=np.float64)
        out[f'{tag}_loss'] = torch.stack([tot, cl, bl]).detach()
        for l in range(5):
            out[f'{tag}_cls_t{l}'] = cls_t[l].to(torch.int16)
            out[f'{tag}_box_t{l}'] = box_t[l]
            if C <= 20:
                out[f'{tag}_gcls{l}'] = co[l].grad
                out[f'{tag}_gbox{l}'] = bo[l].grad
            else:
                gc_, gb_ = co[l].grad.double().flatten(), bo[l].grad.double().flatten()
                wc = torch.arange(gc_.numel(), dtype=torch.float64) % 1009 + 1
                wb = torch.arange(gb_.numel(), dtype=torch.float64) % 1009 + 1
                out[f'{tag}_gcls{l}_chk'] = torch.stack([gc_.sum(), gc_.abs().sum(), (gc_ * wc).sum()])
                out[f'{tag}_gbox{l}_chk'] = torch.stack([gb_.sum(), gb_.abs().sum(), (gb_ * wb).sum()])
                out[f'{tag}_gbox{l}'] = bo[l].grad
        out[f'{tag}_npos'] = npos
    save('loss', **out)


# ----------------------------------------------------------------------------- post-process
def gold_postprocess():
    out = {}
    cases = [
        # tag, size, B, C, K, D, regime, seed
        ('pp128', 128, 3, 20, 500, 100, 'dense', 31),
        ('pp256', 256, 2, 90, 5000, 100, 'dense', 32),
        ('pp256s', 256, 2, 90, 5000, 100, 'sparse', 33),
        ('pp128c1', 128, 2, 1, 300, 30, 'dense', 34),
        ('pp512', 512, 2, 90, 5000, 100, 'sparse', 35),
    ]
    for tag, size, B, C, K, D, regime, seed in cases:
        anc = mk_anchors(size)
        if regime == 'dense':
            co, bo = synth.head_outputs(seed, B, size, C)
        else:
            co, bo = synth.planted_outputs(seed, B, size, C)
        co = [torch.from_numpy(x) for x in co]
        bo = [torch.from_numpy(x) for x in bo]
        cls_k, box_k, idx, klass = _post_process(co, bo, 5, C, K)
        out[f'{tag}_params'] = np.array([size, B, C, K, D, seed, regime == 'sparse'])
        out[f'{tag}_cls'] = cls_k
        out[f'{tag}_box'] = box_k
        out[f'{tag}_idx'] = idx.to(torch.int32)
        out[f'{tag}_klass'] = klass.to(torch.int16)
        # per-anchor OOD rows (the [K, C] gather the reference does at bench.py:51-52)
        allc = torch.cat([c.permute(0, 2, 3, 1).reshape(B, -1, C) for c in co], 1)
        for i in range(B):
            for soft in (False, True):
                for scaled in (False, True):
                    if scaled:
                        scale = torch.tensor(1.0 + 0.25 * i)
                        isz = torch.tensor([size * 1.1, size * 0.9])
                    else:
                        scale, isz = None, None
                    det = generate_detections(cls_k[i], box_k[i], anc.boxes, idx[i], klass[i], scale, isz,
                                              max_det_per_image=D, soft_nms=soft)
                    key = f'{tag}_det_b{i}_{"soft" if soft else "hard"}{"_scaled" if scaled else ""}'
                    out[key] = det
            # energy / max-logit of the rows the first D hard-NMS detections came from is derived in the
            # tests from the logits themselves; store the reference-side gather for the first 8 top-k rows
            rows = allc[i][idx[i][:8]]
            out[f'{tag}_rows_b{i}'] = rows
    save('postprocess', **out)


# ----------------------------------------------------------------------------- soft-nms / decode
def gold_softnms():
    out = {}
    for tag, n, seed, gauss in [('g300', 300, 41, True), ('l300', 300, 42, False), ('g1500', 1500, 43, True)]:
        boxes, scores, classes = synth.nms_candidates(seed, n, 512, 10)
        b, s, c = torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(classes)
        i1, s1 = soft_nms(b, s, method_gaussian=gauss, sigma=0.5, iou_threshold=0.5, score_threshold=0.005)
        i2, s2 = batched_soft_nms(b, s, c, method_gaussian=gauss, sigma=0.5, iou_threshold=0.3, score_threshold=0.001)
        keep = tvb._batched_nms_coordinate_trick(b, s, c, 0.3)
        keep5 = tvb.nms(b, s, 0.5)
        out[f'{tag}_plain_idx'], out[f'{tag}_plain_sc'] = i1.to(torch.int32), s1
        out[f'{tag}_batched_idx'], out[f'{tag}_batched_sc'] = i2.to(torch.int32), s2
        out[f'{tag}_hard_keep'] = keep.to(torch.int32)
        out[f'{tag}_hard_keep_plain'] = keep5.to(torch.int32)
    rs = np.random.RandomState(44)
    anc = mk_anchors(128).boxes
    codes = torch.from_numpy((rs.standard_normal((anc.shape[0], 4)) * 0.3).astype(np.float32))
    out['decode_yxyx'] = decode_box_outputs(codes, anc, output_xyxy=False)
    out['decode_xyxy'] = decode_box_outputs(codes, anc, output_xyxy=True)
    save('softnms', **out)


# ----------------------------------------------------------------------------- evaluation
def gold_evaluation():
    from effdet.evaluation.per_image_evaluation import PerImageEvaluation
    out = {}
    for tag, seed, n_det, n_gt, C, nms_iou, nms_max, flags in synth.EVAL_CASES:
        det, scores, cls, gtb, gtc, dif, gof = synth.eval_case(seed, n_det, n_gt, C)
        if not flags:
            dif[:] = False
            gof[:] = False
        ev = PerImageEvaluation(C, matching_iou_threshold=0.5, nms_iou_threshold=nms_iou, nms_max_output_boxes=nms_max)
        sc, tp, corloc = ev.compute_object_detection_metrics(det.copy(), scores.copy(), cls.copy(), gtb.copy(), gtc.copy(), dif.copy(), gof.copy())
        out[f'{tag}_corloc'] = np.asarray(corloc).astype(np.uint8)
        for c in range(C):
            out[f'{tag}_scores_{c}'] = np.asarray(sc[c], np.float32)
            out[f'{tag}_tp_{c}'] = np.asarray(tp[c]).astype(np.float32)
    save('evaluation', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['anchors', 'labeler', 'loss', 'postprocess', 'softnms', 'evaluation']
    for w in which:
        globals()['gold_' + w]()

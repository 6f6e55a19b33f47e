"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference python path.

The golden files were written by tests/golden/make_golden.py next to /root/reference; nothing
here reads the reference tree.  Bars: bit-exact for assignments / indices / kept sets, 1e-5
relative (stated per assert) for fp32 values.
"""
import numpy as np
import pytest

import synth
from oracle import oracle as orc

RTOL = 1e-5


def anchors_for(size, scale=4.0):
    return orc.anchor_boxes(3, 7, 3, synth.ASPECTS, scale, (size, size))


# ------------------------------------------------------------------ anchors (anchors.py:194-299)
def test_anchor_table(golden):
    g = golden('anchors')
    for name, (size, scale) in synth.MODEL_SHAPES.items():
        b = anchors_for(size, scale)
        assert b.shape[0] == int(g[f'{name}_count']) == synth.num_anchors(size)
        np.testing.assert_array_equal(b[:32], g[f'{name}_head'])
        np.testing.assert_array_equal(b[-32:], g[f'{name}_tail'])
        np.testing.assert_allclose(b.astype(np.float64).sum(0), g[f'{name}_sum64'], rtol=1e-13)
        wgt = (np.arange(b.shape[0], dtype=np.float64) % 1009 + 1)[:, None]
        # float64 checksum over exact fp32 values: only the summation order differs
        np.testing.assert_allclose((b.astype(np.float64) * wgt).sum(0), g[f'{name}_wsum64'], rtol=1e-13)
    np.testing.assert_array_equal(anchors_for(128), g['s128'])
    np.testing.assert_array_equal(orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 3.0, (256, 256)), g['s256_scale3'])
    # survey KATs
    d0 = anchors_for(512)
    np.testing.assert_allclose(d0[0], [-12, -12, 20, 20])
    np.testing.assert_allclose(d0[1], [-7.2, -18.4, 15.2, 26.4], rtol=1e-6)


def test_repo_anchors_match_the_reference_table(golden):
    """The package's own Anchors (the table the kernels and every GPU test use) against the reference's table for
    d0 / d3 / d5 / d7: count, first and last rows bit for bit, and the float64 checksums; plus bit equality with
    the oracle's table and the plane descriptors the gt-centric labeler derives from it."""
    from ood_object_detection_b200.anchors import Anchors
    g = golden('anchors')
    for name, (size, scale) in synth.MODEL_SHAPES.items():
        anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size))
        b = anc.boxes.numpy()
        assert b.dtype == np.float32 and b.shape == (int(g[f'{name}_count']), 4)
        np.testing.assert_array_equal(b[:32], g[f'{name}_head'])
        np.testing.assert_array_equal(b[-32:], g[f'{name}_tail'])
        np.testing.assert_allclose(b.astype(np.float64).sum(0), g[f'{name}_sum64'], rtol=1e-13)
        wgt = (np.arange(b.shape[0], dtype=np.float64) % 1009 + 1)[:, None]
        np.testing.assert_allclose((b.astype(np.float64) * wgt).sum(0), g[f'{name}_wsum64'], rtol=1e-13)
        np.testing.assert_array_equal(b, anchors_for(size, scale))
        # plane descriptors: cell (0, 0) of every (level, shape) grid is the table's own first row of that plane
        desc = anc.plane_desc.numpy()
        assert desc.shape == (5 * 9, 12)
        for k in (0, 8, 9, 44):
            off, a = int(desc[k, 9]), int(desc[k, 10])
            row = b[off + a].astype(np.float64)
            np.testing.assert_allclose([desc[k, 0] - desc[k, 4], desc[k, 1] - desc[k, 5], desc[k, 0] + desc[k, 4], desc[k, 1] + desc[k, 5]],
                                       row, rtol=1e-6, atol=1e-4)
    assert 'anchors.plane_desc' not in Anchors(3, 7, 3, synth.ASPECTS, 4.0, (128, 128)).state_dict() \
        and 'plane_desc' not in Anchors(3, 7, 3, synth.ASPECTS, 4.0, (128, 128)).state_dict()


# ------------------------------------------------------------------ labeler
@pytest.mark.parametrize('tag,kw', [('empty', {}), ('zero_iou', {}), ('identical', {}), ('tiny', {}),
                                    ('padded_float', {}), ('collide', {}), ('nofilter', {'filter_valid': False})])
def test_labeler_kat(golden, tag, kw):
    g = golden('labeler')
    anc = anchors_for(512)
    cls_t, box_t, npos, _, _ = orc.batch_label_anchors(anc, [g[f'kat_{tag}_boxes']], [g[f'kat_{tag}_classes']], **kw)
    pos = np.nonzero(cls_t[0] != -1)[0]
    np.testing.assert_array_equal(pos, g[f'kat_{tag}_pos_idx'])
    np.testing.assert_array_equal(cls_t[0][pos], g[f'kat_{tag}_pos_cls'])
    np.testing.assert_allclose(box_t[0][pos], g[f'kat_{tag}_pos_box'], rtol=RTOL, atol=1e-7)
    np.testing.assert_array_equal(npos, g[f'kat_{tag}_npos'])
    neg = np.ones(cls_t.shape[1], bool)
    neg[pos] = False
    assert (box_t[0][neg] == 0).all()


@pytest.mark.parametrize('tag,size,m,seed,integer', [('r256_m10', 256, 10, 11, False), ('r256_m100', 256, 100, 12, False),
                                                     ('r256_int', 256, 24, 13, True), ('r512_m10', 512, 10, 14, False)])
def test_labeler_random(golden, tag, size, m, seed, integer):
    g = golden('labeler')
    gb, _ = synth.gt_boxes(seed, 3, size, m, 90, integer=integer)
    gc = g[f'{tag}_gc']
    cls_t, box_t, npos, _, _ = orc.batch_label_anchors(anchors_for(size), gb, gc)
    np.testing.assert_array_equal(cls_t, g[f'{tag}_cls'].astype(np.int64))
    np.testing.assert_array_equal(npos, g[f'{tag}_npos'])
    np.testing.assert_allclose(box_t, g[f'{tag}_box'], rtol=RTOL, atol=1e-7)
    assert ((box_t != 0) == (g[f'{tag}_box'] != 0)).all()  # the huber mask keys on target != 0


@pytest.mark.parametrize('thr', [0.4, 0.7])
def test_labeler_thresholds(golden, thr):
    g = golden('labeler')
    gb, gc = synth.gt_boxes(15, 2, 256, 20, 90)
    cls_t, box_t, npos, _, _ = orc.batch_label_anchors(anchors_for(256), gb, gc, match_threshold=thr)
    k = f'thr{int(thr * 10)}'
    np.testing.assert_array_equal(cls_t, g[k + '_cls'].astype(np.int64))
    np.testing.assert_array_equal(npos, g[k + '_npos'])
    np.testing.assert_allclose(box_t, g[k + '_box'], rtol=RTOL, atol=1e-7)


def test_labeler_task_cls(golden):
    g = golden('labeler')
    cls_t, box_t, npos, _, cls_out = orc.batch_label_anchors(anchors_for(256), g['task_boxes'], g['task_classes_in'],
                                                             task_cls=5)
    np.testing.assert_array_equal(cls_out[0], g['task_classes_out'][0])
    np.testing.assert_array_equal(cls_t, g['task_cls'].astype(np.int64))
    np.testing.assert_array_equal(npos, g['task_npos'])
    np.testing.assert_allclose(box_t, g['task_box'], rtol=RTOL, atol=1e-7)


# ------------------------------------------------------------------ loss (loss.py:224-298)
LOSS_TAGS = ['new_c90', 'new_c1', 'new_smooth', 'legacy', 'legacy_g0', 'new_256']


def loss_case(g, tag):
    size, B, C, M, alpha, gamma, delta, w, sm, legacy, plant, s_gt, s_out, s_pl = g[f'{tag}_params']
    size, B, C, M = int(size), int(B), int(C), int(M)
    co, bo = synth.head_outputs(int(s_out), B, size, C, tie_free=False)
    cls_t = [g[f'{tag}_cls_t{l}'].astype(np.int64) for l in range(5)]
    box_t = [g[f'{tag}_box_t{l}'] for l in range(5)]
    return dict(size=size, B=B, C=C, M=M, alpha=float(alpha), gamma=float(gamma), delta=float(delta), w=float(w),
                sm=float(sm), legacy=bool(legacy), co=co, bo=bo, cls_t=cls_t, box_t=box_t, npos=g[f'{tag}_npos'],
                s_gt=int(s_gt), plant=bool(plant), s_pl=int(s_pl))


@pytest.mark.parametrize('tag', LOSS_TAGS)
def test_loss(golden, tag):
    g = golden('loss')
    c = loss_case(g, tag)
    tot, cl, bl, gcs, gbs = orc.loss_fn(c['co'], c['bo'], c['cls_t'], c['box_t'], c['npos'], c['C'], c['alpha'],
                                        c['gamma'], c['delta'], c['w'], c['sm'], c['legacy'], want_grad=True)
    np.testing.assert_allclose([tot, cl, bl], g[f'{tag}_loss'], rtol=RTOL)
    for l in range(5):
        np.testing.assert_allclose(gbs[l], g[f'{tag}_gbox{l}'], rtol=RTOL, atol=1e-9)
        if f'{tag}_gcls{l}' in g:
            np.testing.assert_allclose(gcs[l], g[f'{tag}_gcls{l}'], rtol=1e-4, atol=1e-9)
        else:
            x = gcs[l].astype(np.float64).ravel()
            wgt = np.arange(x.size, dtype=np.float64) % 1009 + 1
            np.testing.assert_allclose([x.sum(), np.abs(x).sum(), (x * wgt).sum()], g[f'{tag}_gcls{l}_chk'], rtol=RTOL)


def test_loss_targets_match_labeler(golden):
    """The stored loss targets are what the (oracle) labeler produces for the same gt seeds."""
    g = golden('loss')
    for tag in ('new_c90', 'new_256'):
        c = loss_case(g, tag)
        gb, gc = synth.gt_boxes(c['s_gt'], c['B'], c['size'], c['M'], c['C'])
        cls_f, box_f, npos, _, _ = orc.batch_label_anchors(anchors_for(c['size']), gb, gc)
        fhw = synth.feat_hw(c['size'])
        for l, (ct, bt) in enumerate(zip(orc.split_levels(cls_f, fhw), orc.split_levels(box_f, fhw))):
            np.testing.assert_array_equal(ct, c['cls_t'][l])
            np.testing.assert_allclose(bt, c['box_t'][l], rtol=RTOL, atol=1e-7)
        np.testing.assert_array_equal(npos, c['npos'])


# ------------------------------------------------------------------ post-process (bench.py:12-56, anchors.py:95-172)
PP_TAGS = ['pp128', 'pp256', 'pp256s', 'pp128c1', 'pp512']


def pp_inputs(g, tag):
    size, B, C, K, D, seed, sparse = [int(v) for v in g[f'{tag}_params']]
    co, bo = (synth.planted_outputs if sparse else synth.head_outputs)(seed, B, size, C)
    return size, B, C, K, D, co, bo


@pytest.mark.parametrize('tag', PP_TAGS)
def test_post_process(golden, tag):
    g = golden('postprocess')
    size, B, C, K, D, co, bo = pp_inputs(g, tag)
    cls_k, box_k, idx, klass = orc.post_process(co, bo, 5, C, K)
    np.testing.assert_array_equal(idx, g[f'{tag}_idx'].astype(np.int64))
    np.testing.assert_array_equal(klass, g[f'{tag}_klass'].astype(np.int64))
    np.testing.assert_array_equal(cls_k, g[f'{tag}_cls'])
    np.testing.assert_array_equal(box_k, g[f'{tag}_box'])
    rows = orc.gather_logit_rows(co, idx[:, :8], C)
    for i in range(B):
        np.testing.assert_array_equal(rows[i], g[f'{tag}_rows_b{i}'])


@pytest.mark.parametrize('tag', PP_TAGS)
@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('scaled', [False, True])
def test_generate_detections(golden, tag, soft, scaled):
    g = golden('postprocess')
    size, B, C, K, D, co, bo = pp_inputs(g, tag)
    anc = anchors_for(size)
    for i in range(B):
        scale = np.float32(1.0 + 0.25 * i) if scaled else None
        isz = np.array([size * 1.1, size * 0.9], np.float32) if scaled else None
        det = orc.generate_detections(g[f'{tag}_cls'][i], g[f'{tag}_box'][i], anc, g[f'{tag}_idx'][i],
                                      g[f'{tag}_klass'][i], scale, isz, D, soft)
        ref = g[f'{tag}_det_b{i}_{"soft" if soft else "hard"}{"_scaled" if scaled else ""}']
        assert det.shape == ref.shape
        np.testing.assert_array_equal(det[:, 5], ref[:, 5])  # same detections in the same order
        np.testing.assert_allclose(det[:, :5], ref[:, :5], rtol=RTOL, atol=1e-6)


# ------------------------------------------------------------------ soft-nms module / decode
@pytest.mark.parametrize('tag,n,seed,gauss', [('g300', 300, 41, True), ('l300', 300, 42, False), ('g1500', 1500, 43, True)])
def test_soft_nms(golden, tag, n, seed, gauss):
    g = golden('softnms')
    boxes, scores, classes = synth.nms_candidates(seed, n, 512, 10)
    i1, s1 = orc.soft_nms(boxes, scores, gauss, 0.5, 0.5, 0.005)
    np.testing.assert_array_equal(i1, g[f'{tag}_plain_idx'].astype(np.int64))
    np.testing.assert_allclose(s1, g[f'{tag}_plain_sc'], rtol=RTOL)
    i2, s2 = orc.batched_soft_nms(boxes, scores, classes, gauss, 0.5, 0.3, 0.001)
    np.testing.assert_array_equal(i2, g[f'{tag}_batched_idx'].astype(np.int64))
    np.testing.assert_allclose(s2, g[f'{tag}_batched_sc'], rtol=RTOL)
    np.testing.assert_array_equal(orc.batched_nms(boxes, scores, classes, 0.3), g[f'{tag}_hard_keep'].astype(np.int64))
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.5), g[f'{tag}_hard_keep_plain'].astype(np.int64))
    # size-independent property: the first D rounds do not depend on later ones
    i3, s3 = orc.soft_nms(boxes, scores, gauss, 0.5, 0.5, 0.005, max_rounds=50)
    np.testing.assert_array_equal(i3, i1[:50])
    np.testing.assert_array_equal(s3, s1[:50])


def test_decode(golden):
    g = golden('softnms')
    rs = np.random.RandomState(44)
    anc = anchors_for(128)
    codes = (rs.standard_normal((anc.shape[0], 4)) * 0.3).astype(np.float32)
    np.testing.assert_allclose(orc.decode_box_outputs(codes, anc, False), g['decode_yxyx'], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(orc.decode_box_outputs(codes, anc, True), g['decode_xyxy'], rtol=RTOL, atol=1e-5)


def test_ood_definition():
    rs = np.random.RandomState(5)
    rows = (rs.standard_normal((64, 90)) * 2 - 4).astype(np.float32)
    e, m = orc.ood_scores(rows)
    ref = -np.log(np.exp(rows.astype(np.float64)).sum(1))
    np.testing.assert_allclose(e, ref, rtol=1e-6)
    np.testing.assert_array_equal(m, rows.max(1))


# ------------------------------------------------------------------ per-image evaluation
def assert_eval_matches_golden(g, tag, C, scores, cls, label, corloc):
    """Per-slot labels (1 tp / 0 fp / -1 ignored / -2 removed) against the reference's per-class (scores,
    tp_fp_labels) arrays -- same (score, label) sequence per class -- and its CorLoc flags."""
    np.testing.assert_array_equal(np.asarray(corloc), g[f'{tag}_corloc'])
    for c in range(C):
        sel = (cls == c) & (label >= 0)
        order = np.argsort(-scores[sel], kind='stable')
        np.testing.assert_array_equal(scores[sel][order], g[f'{tag}_scores_{c}'])
        np.testing.assert_array_equal(label[sel][order].astype(np.float32), g[f'{tag}_tp_{c}'])


@pytest.mark.parametrize('case', synth.EVAL_CASES, ids=[c[0] for c in synth.EVAL_CASES])
def test_match_detections_golden(golden, case):
    """oracle.match_detections vs PerImageEvaluation.compute_object_detection_metrics of the reference
    (per_image_evaluation.py:29-92) on detections scattered around the gt boxes: plain, with difficult /
    group-of boxes, with the class's own NMS (0.3 / 50), a single class, classes without gt."""
    tag, seed, n_det, n_gt, C, nms_iou, nms_max, flags = case
    g = golden('evaluation')
    det, scores, cls, gtb, gtc, dif, gof = synth.eval_case(seed, n_det, n_gt, C)
    if not flags:
        dif[:] = False
        gof[:] = False
    label, corloc = orc.match_detections(det, scores, cls, gtb, gtc, C, dif, gof, 0.5, nms_iou, nms_max)
    assert_eval_matches_golden(g, tag, C, scores, cls, label, corloc)
    assert (label == 1).sum() > 0 and (label == 0).sum() > 0


def test_anchor_generator_reproduces_the_table():
    """`Anchors.plane_gen` (float64 centre origin, pitch and half sizes per (level, shape) grid) is what the labeler
    kernel recomputes anchors from instead of gathering them: float32(cy0 + y*sy -/+ half) must equal the table bit
    for bit, for every anchor of every model shape (the same IEEE operations in numpy as in the kernel)."""
    from ood_object_detection_b200.anchors import Anchors
    for name, (size, scale) in synth.MODEL_SHAPES.items():
        anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size))
        gen, desc, na = anc.plane_gen.numpy(), anc.plane_desc.numpy(), anc.get_anchors_per_location()
        assert gen.dtype == np.float64 and gen.shape == (desc.shape[0], 6)
        table = anc.boxes.numpy()
        for k in range(gen.shape[0]):
            W, H, off, a = int(desc[k, 7]), int(desc[k, 8]), int(desc[k, 9]), int(desc[k, 10])
            cy0, cx0, sy, sx, hy, hx = gen[k]
            yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
            cy, cx = cy0 + yy.ravel() * sy, cx0 + xx.ravel() * sx
            got = np.stack([cy - hy, cx - hx, cy + hy, cx + hx], 1).astype(np.float32)
            rows = off + np.arange(H * W) * na + a
            np.testing.assert_array_equal(got, table[rows], err_msg=f'{name} plane {k}')

"""GPU parity: odk_topk / odk_detect / odk_soft_nms / odk_nms / odk_ood (through the effdet-API shims
and the C ABI) against the reference goldens and the CPU oracle.  Bars: top-k indices, classes and
kept-detection order bit-exact; boxes / scores / soft scores / OOD scores within 1e-5 relative."""
import numpy as np
import pytest
import torch

import synth
from oracle import oracle as orc
from test_oracle_golden import PP_TAGS, pp_inputs

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DEV = 'cuda:0'


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def assert_dets_close(det, ref, rtol=RTOL):
    """Rows [x0, y0, x1, y1, score, class]: class exact; score within rtol; box coordinates within
    rtol RELATIVE TO THE BOX'S OWN SCALE (its largest |coordinate|): a corner is centre -/+ half-size,
    so a 1-ulp difference between two exp() implementations survives the cancellation as an
    absolute error of ~1 ulp of the larger operand even when the corner itself is near 0."""
    det, ref = np.asarray(det), np.asarray(ref)
    assert det.shape == ref.shape
    np.testing.assert_array_equal(det[:, 5], ref[:, 5])
    np.testing.assert_allclose(det[:, 4], ref[:, 4], rtol=rtol)
    scale = np.maximum(np.abs(ref[:, :4]).max(axis=1, keepdims=True), 1.0)
    err = np.abs(det[:, :4] - ref[:, :4]) / scale
    assert err.max(initial=0.0) <= rtol, f'box error {err.max()} relative to box scale exceeds {rtol}'


def anchors_t(size, scale=4.0):
    return t(orc.anchor_boxes(3, 7, 3, synth.ASPECTS, scale, (size, size)))


@pytest.mark.parametrize('tag', PP_TAGS)
def test_post_process_golden(golden, tag):
    from ood_object_detection_b200.bench import _post_process
    g = golden('postprocess')
    size, B, C, K, D, co, bo = pp_inputs(g, tag)
    cls_k, box_k, idx, klass = _post_process([t(x) for x in co], [t(x) for x in bo], 5, C, K)
    assert cls_k.shape == (B, K, 1) and idx.dtype == torch.int64 and klass.dtype == torch.int64
    np.testing.assert_array_equal(idx.cpu().numpy(), g[f'{tag}_idx'].astype(np.int64))
    np.testing.assert_array_equal(klass.cpu().numpy(), g[f'{tag}_klass'].astype(np.int64))
    np.testing.assert_array_equal(cls_k.cpu().numpy(), g[f'{tag}_cls'])
    np.testing.assert_array_equal(box_k.cpu().numpy(), g[f'{tag}_box'])


@pytest.mark.parametrize('name,B,C,K,sparse', [('d0', 4, 90, 5000, False), ('d0', 2, 90, 5000, True),
                                               ('d3', 2, 90, 5000, False), ('d0', 3, 1, 2000, False),
                                               ('d0', 1, 400, 5000, False)])
def test_post_process_vs_oracle(name, B, C, K, sparse):
    from ood_object_detection_b200.bench import _post_process
    size, _ = synth.MODEL_SHAPES[name]
    co, bo = (synth.planted_outputs if sparse else synth.head_outputs)(600 + B + C, B, size, C)
    ref = orc.post_process(co, bo, 5, C, K)
    got = _post_process([t(x) for x in co], [t(x) for x in bo], 5, C, K)
    for r, g_ in zip(ref, got):
        np.testing.assert_array_equal(g_.cpu().numpy(), r)


def test_post_process_degenerate_inputs():
    """Inputs the sampled threshold cannot handle go through the exact radix-select path:
    all-equal logits (pure index ties), heavily quantised logits, one huge outlier cluster."""
    from ood_object_detection_b200.bench import _post_process
    size, B, C, K = 256, 4, 20, 3000
    co, bo = synth.head_outputs(71, B, size, C, tie_free=False)
    for c in co:
        c[0] = -4.5                                            # constant image
        c[1] = np.round(c[1] * 2) / 2                          # ~20 distinct values
        c[2] = np.where(c[2] > -4.6, np.float32(1.25), c[2])   # half the image tied at the top
    ref = orc.post_process(co, bo, 5, C, K)
    got = _post_process([t(x) for x in co], [t(x) for x in bo], 5, C, K)
    for r, g_ in zip(ref, got):
        np.testing.assert_array_equal(g_.cpu().numpy(), r)
    # sortedness / membership properties on the constant image: first K flat indices
    flat = got[2][0].cpu().numpy() * C + got[3][0].cpu().numpy()
    np.testing.assert_array_equal(flat, np.arange(K))


@pytest.mark.parametrize('case', ['sign_change', 'many_positives', 'outlier', 'inf_outlier', 'coarse_ties', 'narrow'])
def test_post_process_select_paths(case):
    """The select kernel's counting sort (sub-bins on the key bits, value-linear second level for the top
    sub-bin) and its bitonic fallback: score distributions that stress each branch, against the oracle
    (ties: ascending flat index, torch.topk's order on the reference's CPU path)."""
    from ood_object_detection_b200.bench import _post_process
    size, B, C, K = 256, 3, 20, 3000
    mean, std = {'sign_change': (-4.0, 1.5), 'many_positives': (-1.0, 2.0), 'narrow': (-4.6, 0.01)}.get(case, (-4.6, 1.5))
    co, bo = synth.head_outputs(90 + len(case), B, size, C, mean=mean, std=std, tie_free=(case not in ('coarse_ties', 'narrow')))
    if case == 'outlier':
        co[0][:, 3, 1, 1] = 1.0e30          # one huge score stretches the value range of the top sub-bin
    if case == 'inf_outlier':
        co[1][:, 5, 0, 2] = np.inf
    if case == 'coarse_ties':
        for c in co:
            c[...] = np.round(c * 16) / 16   # hundreds of equal values per step: big bins, index-order ties
    ref = orc.post_process(co, bo, 5, C, K)
    got = _post_process([t(x) for x in co], [t(x) for x in bo], 5, C, K)
    for r, g_ in zip(ref, got):
        np.testing.assert_array_equal(g_.cpu().numpy(), r)


def test_post_process_errors():
    from ood_object_detection_b200.bench import _post_process
    co, bo = synth.head_outputs(1, 1, 128, 1)
    with pytest.raises(RuntimeError):
        _post_process([t(x) for x in co], [t(x) for x in bo], 5, 1, 5000)   # k > A*C like torch.topk
    with pytest.raises(RuntimeError):
        _post_process([torch.from_numpy(x) for x in co], [torch.from_numpy(x) for x in bo], 5, 1, 100)   # CPU tensors


@pytest.mark.parametrize('tag', PP_TAGS)
@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('scaled', [False, True])
def test_generate_detections_golden(golden, tag, soft, scaled):
    from ood_object_detection_b200.anchors import generate_detections
    g = golden('postprocess')
    size, B, C, K, D, co, bo = pp_inputs(g, tag)
    anc = anchors_t(size)
    for i in range(B):
        scale = t(np.float32(1.0 + 0.25 * i)) if scaled else None
        isz = t(np.array([size * 1.1, size * 0.9], np.float32)) if scaled else None
        det = generate_detections(t(g[f'{tag}_cls'][i]), t(g[f'{tag}_box'][i]), anc, t(g[f'{tag}_idx'][i].astype(np.int64)),
                                  t(g[f'{tag}_klass'][i].astype(np.int64)), scale, isz, max_det_per_image=D, soft_nms=soft)
        ref = g[f'{tag}_det_b{i}_{"soft" if soft else "hard"}{"_scaled" if scaled else ""}']
        assert_dets_close(det.cpu().numpy(), ref)


@pytest.mark.parametrize('soft', [False, True])
def test_batch_detection_full_chain_vs_oracle(soft):
    """D0, B=4: _post_process -> _batch_detection (one launch) == oracle per-image chain; kept
    source positions bit-exact."""
    from ood_object_detection_b200.anchors import detect_batch
    from ood_object_detection_b200.bench import _post_process, _batch_detection
    size, B, C, K, D = 512, 4, 90, 5000, 100
    co, bo = synth.planted_outputs(81, B, size, C)
    anc_np = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    cls_k, box_k, idx, klass = _post_process([t(x) for x in co], [t(x) for x in bo], 5, C, K)
    dets, count, src = detect_batch(cls_k, box_k, t(anc_np), idx, klass, None, None, D, soft)
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, K)
    for i in range(B):
        ref, rsrc = orc.generate_detections(o_cls[i], o_box[i], anc_np, o_idx[i], o_klass[i], None, None, D, soft,
                                            return_src=True)
        n = int(count[i].item())
        assert n == ref.shape[0]
        np.testing.assert_array_equal(src[i, :n].cpu().numpy(), rsrc)
        assert_dets_close(dets[i, :n].cpu().numpy(), ref)
        assert (dets[i, n:] == 0).all() and (src[i, n:] == -1).all()
    if int(count.min().item()) < D:
        with pytest.raises(RuntimeError):
            _batch_detection(B, cls_k, box_k, t(anc_np), idx, klass, max_det_per_image=D, soft_nms=soft)
    padded = _batch_detection(B, cls_k, box_k, t(anc_np), idx, klass, max_det_per_image=D, soft_nms=soft, pad=True)
    assert padded.shape == (B, D, 6)


def test_generate_detections_unsorted_and_empty():
    """API inputs need not be sorted (torchvision sorts by score itself) and may all fail the filter."""
    from ood_object_detection_b200.anchors import generate_detections
    size, N, C = 256, 700, 12
    anc_np = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    rs = np.random.RandomState(5)
    idx = rs.randint(0, anc_np.shape[0], N).astype(np.int64)
    klass = rs.randint(0, C, N).astype(np.int64)
    cls = (rs.standard_normal((N, 1)) * 2.0 - 2.0).astype(np.float32)
    box = (rs.standard_normal((N, 4)) * 0.3).astype(np.float32)
    for soft in (False, True):
        ref = orc.generate_detections(cls, box, anc_np, idx, klass, None, None, 50, soft)
        got = generate_detections(t(cls), t(box), t(anc_np), t(idx), t(klass), None, None, 50, soft).cpu().numpy()
        assert_dets_close(got, ref)
        none = generate_detections(t(cls - 20), t(box), t(anc_np), t(idx), t(klass), None, None, 50, soft)
        assert none.shape == (0, 6)


@pytest.mark.parametrize('tag,n,seed,gauss', [('g300', 300, 41, True), ('l300', 300, 42, False), ('g1500', 1500, 43, True)])
def test_soft_nms_module(golden, tag, n, seed, gauss):
    from ood_object_detection_b200 import soft_nms as S
    g = golden('softnms')
    boxes, scores, classes = synth.nms_candidates(seed, n, 512, 10)
    i1, s1 = S.soft_nms(t(boxes), t(scores), gauss, 0.5, 0.5, 0.005)
    np.testing.assert_array_equal(i1.cpu().numpy(), g[f'{tag}_plain_idx'].astype(np.int64))
    np.testing.assert_allclose(s1.cpu().numpy(), g[f'{tag}_plain_sc'], rtol=RTOL)
    i2, s2 = S.batched_soft_nms(t(boxes), t(scores), t(classes), gauss, 0.5, 0.3, 0.001)
    np.testing.assert_array_equal(i2.cpu().numpy(), g[f'{tag}_batched_idx'].astype(np.int64))
    np.testing.assert_allclose(s2.cpu().numpy(), g[f'{tag}_batched_sc'], rtol=RTOL)
    np.testing.assert_array_equal(S.batched_nms(t(boxes), t(scores), t(classes), 0.3).cpu().numpy(),
                                  g[f'{tag}_hard_keep'].astype(np.int64))
    np.testing.assert_array_equal(S.nms(t(boxes), t(scores), 0.5).cpu().numpy(),
                                  g[f'{tag}_hard_keep_plain'].astype(np.int64))
    e_i, e_s = S.batched_soft_nms(t(boxes[:0]), t(scores[:0]), t(classes[:0]))
    assert e_i.numel() == 0 and e_s.numel() == 0


def test_nms_large_vs_oracle():
    from ood_object_detection_b200 import soft_nms as S
    boxes, scores, classes = synth.nms_candidates(9, 5000, 896, 90, n_clusters=300)
    np.testing.assert_array_equal(S.batched_nms(t(boxes), t(scores), t(classes), 0.3).cpu().numpy(),
                                  orc.batched_nms(boxes, scores, classes, 0.3))
    i, s = S.batched_soft_nms(t(boxes), t(scores), t(classes), True, 0.5, 0.3, 0.001)
    oi, os_ = orc.batched_soft_nms(boxes, scores, classes, True, 0.5, 0.3, 0.001)
    np.testing.assert_array_equal(i.cpu().numpy(), oi)
    np.testing.assert_allclose(s.cpu().numpy(), os_, rtol=RTOL)


def test_ood_scores_and_fused_entry():
    from ood_object_detection_b200.bench import detect_with_ood
    size, B, C, K, D = 256, 3, 30, 2000, 40
    co, bo = synth.planted_outputs(91, B, size, C, n_obj=20)
    anc_np = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    out = detect_with_ood([t(x) for x in co], [t(x) for x in bo], t(anc_np), 5, C, K, D, False, temperature=1.0)
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, K)
    for i in range(B):
        ref, rsrc = orc.generate_detections(o_cls[i], o_box[i], anc_np, o_idx[i], o_klass[i], None, None, D, False,
                                            return_src=True)
        n = int(out['count'][i].item())
        assert n == ref.shape[0]
        assert_dets_close(out['detections'][i, :n].cpu().numpy(), ref)
        anchors_i = o_idx[i][rsrc]
        np.testing.assert_array_equal(out['anchor'][i, :n].cpu().numpy(), anchors_i)
        rows = orc.gather_logit_rows(co, anchors_i[None].repeat(B, 0), C)[i]
        e, m = orc.ood_scores(rows, 1.0)
        np.testing.assert_allclose(out['energy'][i, :n].cpu().numpy(), e, rtol=RTOL)
        np.testing.assert_array_equal(out['max_logit'][i, :n].cpu().numpy(), m)
        assert (out['energy'][i, n:] == 0).all()
    # temperature and the torch definition
    from ood_object_detection_b200.ood import ood_scores
    e2, _ = ood_scores([t(x) for x in co], out['anchor'], 5, C, temperature=2.5)
    allc = torch.cat([t(c).permute(0, 2, 3, 1).reshape(B, -1, C) for c in co], 1)
    rows = torch.gather(allc, 1, out['anchor'].clamp(min=0)[:, :, None].expand(-1, -1, C))
    ref = -2.5 * torch.logsumexp(rows / 2.5, dim=2)
    ok = out['anchor'] >= 0
    np.testing.assert_allclose(e2[ok].cpu().numpy(), ref[ok].cpu().numpy(), rtol=RTOL)


# ---- odk_postprocess: the pipelined chain (sample -> persistent stream + per-image tails) ----------------
def _fused_vs_chain(co, bo, size, C, K, D, soft, scale=None, isz=None, anchor_scale=4.0, ood=False, pipeline='staged'):
    """post_process_detect must equal (a) the separate odk_topk -> odk_detect launches bit for bit and
    (b) the oracle chain under the usual bars."""
    from ood_object_detection_b200.anchors import detect_batch
    from ood_object_detection_b200.bench import _post_process, post_process_detect
    B = co[0].shape[0]
    anc_np = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, anchor_scale, (size, size))
    tc, tb, ta = [t(x) for x in co], [t(x) for x in bo], t(anc_np)
    ts = None if scale is None else t(scale)
    tz = None if isz is None else t(isz)
    out = post_process_detect(tc, tb, ta, 5, C, K, D, soft, ts, tz, with_ood=ood, return_topk=True, pipeline=pipeline)
    cls_k, box_k, idx, klass = _post_process(tc, tb, 5, C, K)
    for name, ref in (('cls', cls_k), ('box', box_k), ('indices', idx), ('classes', klass)):
        assert torch.equal(out[name], ref), name
    dets, count, src = detect_batch(cls_k, box_k, ta, idx, klass, ts, tz, D, soft)
    assert torch.equal(out['count'], count)
    assert torch.equal(out['src'], src)
    assert torch.equal(out['detections'], dets)
    anchor = torch.where(src >= 0, torch.gather(idx, 1, src.clamp(min=0).long()), torch.full_like(src, -1).long())
    assert torch.equal(out['anchor'], anchor)
    # without the top-k outputs the detections must not change
    lean = post_process_detect(tc, tb, ta, 5, C, K, D, soft, ts, tz, pipeline=pipeline)
    assert torch.equal(lean['detections'], dets) and torch.equal(lean['count'], count)
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, K)
    for i in range(B):
        ref, rsrc = orc.generate_detections(o_cls[i], o_box[i], anc_np, o_idx[i], o_klass[i],
                                            None if scale is None else scale[i], None if isz is None else isz[i], D, soft,
                                            return_src=True)
        n = int(out['count'][i].item())
        assert n == ref.shape[0]
        np.testing.assert_array_equal(out['src'][i, :n].cpu().numpy(), rsrc)
        assert_dets_close(out['detections'][i, :n].cpu().numpy(), ref)
        assert (out['detections'][i, n:] == 0).all() and (out['src'][i, n:] == -1).all()
        if ood:
            rows = orc.gather_logit_rows(co, o_idx[i][rsrc][None].repeat(B, 0), C)[i]
            e, m = orc.ood_scores(rows, 1.0)
            np.testing.assert_allclose(out['energy'][i, :n].cpu().numpy(), e, rtol=RTOL)
            np.testing.assert_array_equal(out['max_logit'][i, :n].cpu().numpy(), m)
            assert (out['energy'][i, n:] == 0).all()
    return out


PIPELINES = ['staged', 'persistent']


@pytest.mark.parametrize('pipeline', PIPELINES)
@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('sparse', [False, True])
def test_fused_postprocess_d0(soft, sparse, pipeline):
    size, B, C, K, D = 512, 4, 90, 5000, 100
    co, bo = (synth.planted_outputs if sparse else synth.head_outputs)(310 + sparse, B, size, C)
    _fused_vs_chain(co, bo, size, C, K, D, soft, ood=True, pipeline=pipeline)


@pytest.mark.parametrize('pipeline', PIPELINES)
@pytest.mark.parametrize('soft', [False, True])
def test_fused_postprocess_d3_odd_level_scaled(soft, pipeline):
    """D3's 7x7 level: blocks of odd images start 8 bytes off the 16-byte grid (partial first / last groups);
    with img_scale / img_size the boxes are clipped and rescaled."""
    size, B, C, K, D = 896, 3, 90, 5000, 100
    co, bo = synth.head_outputs(320, B, size, C)
    scale = np.array([1.0, 1.25, 1.5], np.float32)
    isz = np.array([[size * 1.1, size * 0.9]] * B, np.float32)
    _fused_vs_chain(co, bo, size, C, K, D, soft, scale, isz, pipeline=pipeline)


@pytest.mark.parametrize('name,B,C,K,D', [('d0', 3, 1, 2000, 30), ('d0', 1, 400, 5000, 100), ('d0', 5, 7, 6144, 64),
                                          ('d0', 2, 90, 300, 100)])
@pytest.mark.parametrize('pipeline', PIPELINES)
def test_fused_postprocess_shapes(name, B, C, K, D, pipeline):
    size, _ = synth.MODEL_SHAPES[name]
    co, bo = synth.head_outputs(330 + B + C, B, size, C)
    _fused_vs_chain(co, bo, size, C, K, D, False, pipeline=pipeline)
    _fused_vs_chain(co, bo, size, C, K, D, True, pipeline=pipeline)


@pytest.mark.parametrize('pipeline', PIPELINES)
def test_fused_postprocess_small_and_degenerate(pipeline):
    """Tiny pyramids (fewer tasks than warps), images that leave the sampled path (constant, quantised, a
    huge tied plateau -> flagged, exact radix select + stand-alone detect behind the same entry point) next
    to ordinary ones in one batch."""
    size, B, C, K, D = 256, 5, 20, 3000, 50
    co, bo = synth.head_outputs(71, B, size, C, tie_free=False)
    for c in co:
        c[0] = -4.5
        c[1] = np.round(c[1] * 2) / 2
        c[2] = np.where(c[2] > -4.6, np.float32(1.25), c[2])
    synth.make_tie_free([c[3:] for c in co])
    out = _fused_vs_chain(co, bo, size, C, K, D, False, ood=True, pipeline=pipeline)
    _fused_vs_chain(co, bo, size, C, K, D, True, pipeline=pipeline)
    assert out['count'].shape == (B,)
    co, bo = synth.head_outputs(72, 2, 128, 3)
    _fused_vs_chain(co, bo, 128, 3, 200, 20, False, pipeline=pipeline)
    _fused_vs_chain(co, bo, 128, 3, 1000, 20, True, pipeline=pipeline)


def test_fused_postprocess_above_register_budget_uses_chain():
    from ood_object_detection_b200.bench import post_process_detect, FUSED_MAX_K
    size, B, C = 512, 2, 90
    co, bo = synth.head_outputs(340, B, size, C)
    anc = anchors_t(size)
    out = post_process_detect([t(x) for x in co], [t(x) for x in bo], anc, 5, C, FUSED_MAX_K + 1000, 100, False)
    assert out['detections'].shape == (B, 100, 6) and int(out['count'].min()) > 0


@pytest.mark.parametrize('pipeline', PIPELINES)
def test_channels_last_head_outputs_are_read_in_place(pipeline):
    """head outputs in channels_last memory format ([B, H, W, C] in memory, what an AMP /
    channels_last head writes, efficientdet.py:405-414) go through _post_process, odk_postprocess and the OOD
    scores WITHOUT a layout copy and give bit-identical results; mixed layouts per level too (D3: its 7x7 level
    has an odd plane size)."""
    from ood_object_detection_b200 import _lib
    from ood_object_detection_b200.bench import _post_process, post_process_detect
    size, B, C, K, D = 896, 2, 90, 5000, 100
    co, bo = synth.planted_outputs(350, B, size, C)
    anc = anchors_t(size)
    tc, tb = [t(x) for x in co], [t(x) for x in bo]
    cl = [x.contiguous(memory_format=torch.channels_last) for x in tc]
    bl = [x.contiguous(memory_format=torch.channels_last) for x in tb]
    assert _lib.prep_levels(cl, 5)[1] == 0b11111 and all(a.data_ptr() == b_.data_ptr() for a, b_ in zip(_lib.prep_levels(cl, 5)[0], cl))
    ref = _post_process(tc, tb, 5, C, K)
    for cls_in, box_in in ((cl, bl), (cl, tb), ([cl[0], tc[1], cl[2], tc[3], cl[4]], [tb[0], bl[1], tb[2], bl[3], tb[4]])):
        got = _post_process(cls_in, box_in, 5, C, K)
        for r, g_ in zip(ref, got):
            assert torch.equal(r, g_)
    for soft in (False, True):
        want = post_process_detect(tc, tb, anc, 5, C, K, D, soft, with_ood=True, pipeline=pipeline)
        got = post_process_detect(cl, bl, anc, 5, C, K, D, soft, with_ood=True, pipeline=pipeline)
        for key in ('detections', 'count', 'src', 'anchor', 'max_logit'):
            assert torch.equal(want[key], got[key]), key
        np.testing.assert_allclose(got['energy'].cpu().numpy(), want['energy'].cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize('pipeline', PIPELINES)
def test_fused_postprocess_more_images_than_sms(pipeline):
    """B = 180 images (> 148 SMs): the tail kernel needs more than one wave, the collect grid has a single CTA per
    image, the persistent pipeline queues tails.  Fused entry point == the stage chain (itself checked against the
    oracle above), bit for bit, hard and soft."""
    from ood_object_detection_b200.anchors import detect_batch
    from ood_object_detection_b200.bench import _post_process, post_process_detect
    size, B, C, K, D = 128, 180, 8, 500, 20
    co, bo = synth.head_outputs(350, B, size, C)
    tc, tb, ta = [t(x) for x in co], [t(x) for x in bo], anchors_t(size)
    cls_k, box_k, idx, klass = _post_process(tc, tb, 5, C, K)
    for soft in (False, True):
        dets, count, src = detect_batch(cls_k, box_k, ta, idx, klass, None, None, D, soft)
        out = post_process_detect(tc, tb, ta, 5, C, K, D, soft, with_ood=True, pipeline=pipeline)
        assert torch.equal(out['count'], count) and torch.equal(out['src'], src) and torch.equal(out['detections'], dets)
        assert int(count.min()) > 0

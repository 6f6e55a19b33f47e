"""SURVEY 8f rows 2-3 on the GPU: device-side collate labelling and the evaluator hand-off."""
import numpy as np
import pytest
import torch

import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_label_batch_targets_matches_collate_semantics():
    """DetectionFastCollate calls label_anchors(filter_valid=False) on -1 padded [100,4]/[100] gt."""
    from ood_object_detection_b200.anchors import Anchors, AnchorLabeler
    from ood_object_detection_b200.pipeline import label_batch_targets
    size, B, C = 256, 3, 30
    anc = Anchors(3, 7, 3, synth.ASPECTS, 4.0, (size, size)).to(DEV)
    lab = AnchorLabeler(anc, C)
    gb, gc = synth.gt_boxes(21, B, size, 7, C)
    bbox = -np.ones((B, 100, 4), np.float32)
    cls = -np.ones((B, 100), np.float32)          # the loader's float class tensor
    bbox[:, :7], cls[:, :7] = gb, gc
    target = label_batch_targets(lab, {'bbox': torch.from_numpy(bbox).to(DEV), 'cls': torch.from_numpy(cls).to(DEV)})
    oc, ob, onp, _, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), list(bbox), list(cls), filter_valid=False)
    fhw = synth.feat_hw(size)
    for l, (ct, bt) in enumerate(zip(orc.split_levels(oc, fhw), orc.split_levels(ob, fhw))):
        assert target[f'label_cls_{l}'].dtype == torch.int64
        np.testing.assert_array_equal(target[f'label_cls_{l}'].cpu().numpy(), ct)
        np.testing.assert_allclose(target[f'label_bbox_{l}'].cpu().numpy(), bt, rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(target['label_num_positives'].cpu().numpy(), onp)
    assert (oc == -2).any()    # the padded rows really produce ignore targets, as in the reference
    # and DetBenchTrain(create_labeler=False)-style consumption: loss on these targets
    from ood_object_detection_b200.loss import loss_fn
    co, bo = synth.head_outputs(22, B, size, C, tie_free=False)
    out = loss_fn([torch.from_numpy(x).to(DEV) for x in co], [torch.from_numpy(x).to(DEV) for x in bo],
                  [target[f'label_cls_{l}'] for l in range(5)], [target[f'label_bbox_{l}'] for l in range(5)],
                  target['label_num_positives'], C, 0.25, 1.5, 0.1, 50.0)
    ref = orc.loss_fn(co, bo, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0)
    np.testing.assert_allclose([float(v) for v in out], ref, rtol=1e-5)


def test_detections_for_evaluator():
    from ood_object_detection_b200.pipeline import detections_for_evaluator
    rs = np.random.RandomState(3)
    dets = np.zeros((4, 10, 6), np.float32)
    count = np.array([10, 3, 0, 7], np.int32)
    for i, n in enumerate(count):
        dets[i, :n] = rs.uniform(0, 100, (n, 6))
    out = detections_for_evaluator(torch.from_numpy(dets).to(DEV), torch.from_numpy(count).to(DEV))
    for i, n in enumerate(count):
        assert out[i]['bbox'].shape == (n, 4)
        np.testing.assert_array_equal(out[i]['bbox'], dets[i, :n][:, [1, 0, 3, 2]])   # pretrain.py:248
        np.testing.assert_array_equal(out[i]['scores'], dets[i, :n, 4])
        np.testing.assert_array_equal(out[i]['cls'], dets[i, :n, 5])
    raw = detections_for_evaluator(torch.from_numpy(dets).to(DEV), torch.from_numpy(count).to(DEV), yxyx=False)
    np.testing.assert_array_equal(raw[0]['bbox'], dets[0, :10, :4])


@pytest.mark.parametrize('grid', [True, False])
def test_partial_sums_buffer_equals_fused_loss(grid):
    """distributed.local_partial_sums: both kernels write one 4-float buffer; normalising it (what the
    all-reduce path does on every rank) gives the oracle's loss."""
    from ood_object_detection_b200.anchors import Anchors, AnchorLabeler
    from ood_object_detection_b200 import distributed as D
    size, B, C, M = 256, 4, 20, 6
    anc = Anchors(3, 7, 3, synth.ASPECTS, 4.0, (size, size)).to(DEV)
    lab = AnchorLabeler(anc, C)
    lab.use_grid_kernel = grid
    gb, gc = synth.gt_boxes(31, B, size, M, C)
    co, bo = synth.head_outputs(32, B, size, C, tie_free=False)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)
    unit = torch.ones((1,), device=DEV)
    buf = D.local_partial_sums(lab, [torch.from_numpy(x).to(DEV) for x in co], [torch.from_numpy(x).to(DEV) for x in bo],
                               torch.from_numpy(gb).to(DEV), torch.from_numpy(gc).to(DEV), unit, **kw)
    oc, ob, onp, _, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), list(gb), list(gc))
    fhw = synth.feat_hw(size)
    ref = orc.loss_fn(co, bo, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0)
    assert float(buf[3]) == float(onp.sum()) + 1.0
    got = D.all_reduce_partial_sums(buf)      # single process: just the normalisation
    np.testing.assert_allclose([float(v) for v in got], ref, rtol=1e-5)


def test_peer_mailbox_single_rank_sequences():
    """odk_partials_publish / odk_partials_collect with world = 1 (the mailbox is local memory): the
    alternating slots, the device-side sequence counters and the normalisation (loss.py:261,297)."""
    from ood_object_detection_b200.distributed import PeerMailbox
    mb = PeerMailbox(DEV, local_only=True)
    for step in range(5):
        vals = torch.tensor([10.0 + step, 4.0 + step, 0.12, 7.0 + step], device=DEV)   # slot 3 = sum(num_pos) + 1
        mb.publish(vals)
        (tot, cl, bx), status = mb.collect()
        assert int(status) == 0
        n = 7.0 + step
        np.testing.assert_allclose([float(tot), float(cl), float(bx)], [(10.0 + step) / n, (4.0 + step) / n, 0.12 / n], rtol=1e-6)
    # a collect with nothing published gives up after timeout_ms of wall clock and says so instead of hanging the GPU:
    # NaN losses, a STICKY status (1 + the sequence number that failed), and the record set is waited for again
    mb.timeout_ms = 20
    (tot, _, _), status = mb.collect()
    assert int(status) == 1 + 6 and torch.isnan(tot)
    with pytest.raises(RuntimeError, match='did not arrive'):
        mb.check()
    mb.publish(torch.tensor([3.0, 2.0, 1.0, 2.0], device=DEV))
    (tot, cl, bx), status = mb.collect()          # the same sequence number, now present
    np.testing.assert_allclose([float(tot), float(cl), float(bx)], [1.5, 1.0, 0.5], rtol=1e-6)
    assert int(status) == 1 + 6                   # sticky: the caller decides when to look and clear


def test_peer_mailbox_late_rank_is_never_summed_stale():
    """Two ranks emulated on one GPU (rank 1 publishes into rank 0's mailbox from its own counters): when rank 1 is
    late, rank 0's collect times out WITHOUT consuming the step -- it must not sum whatever older record sits in
    rank 1's slot -- and picks the step up once the record has arrived."""
    from ood_object_detection_b200 import _lib
    lib = _lib.lib()
    nbytes = int(lib.odk_mailbox_bytes(2))
    mb0 = torch.zeros((nbytes,), dtype=torch.uint8, device=DEV)
    mb1 = torch.zeros((nbytes,), dtype=torch.uint8, device=DEV)
    ptrs = (_lib.ctypes.c_void_p * 2)(mb0.data_ptr(), mb1.data_ptr())
    out, status = torch.zeros(3, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    st = _lib.stream_ptr(torch.device(DEV))

    def publish(rank, vals):
        v = torch.tensor(vals, dtype=torch.float32, device=DEV)
        _lib.check(lib.odk_partials_publish(_lib.ptr(v), ptrs, 2, rank, st))
        torch.cuda.synchronize()

    def collect(timeout_ms):
        _lib.check(lib.odk_partials_collect(_lib.ptr(mb0), 2, _lib.ptr(out), _lib.ptr(status), timeout_ms, st))
        torch.cuda.synchronize()
        return out.cpu().numpy().copy(), int(status)

    publish(0, [4.0, 2.0, 1.0, 3.0]); publish(1, [6.0, 2.0, 3.0, 4.0])       # step 1, both ranks: n = 3 + 4 - 1 = 6
    got, stt = collect(1000)
    np.testing.assert_allclose(got, [10.0 / 6, 4.0 / 6, 4.0 / 6], rtol=1e-6)
    assert stt == 0
    publish(0, [1.0, 1.0, 1.0, 2.0]); publish(1, [1.0, 1.0, 1.0, 2.0])       # step 2
    collect(1000)
    publish(0, [8.0, 4.0, 2.0, 2.0])                                          # step 3: rank 1 is late ...
    got, stt = collect(20)                                                    # ... its slot still holds step 1's record
    assert np.isnan(got).all() and stt == 1 + 3
    publish(1, [4.0, 2.0, 1.0, 3.0])                                          # ... now it arrives
    got, stt = collect(1000)
    np.testing.assert_allclose(got, [12.0 / 4, 6.0 / 4, 3.0 / 4], rtol=1e-6)  # step 3's own records, n = 2 + 3 - 1
    assert stt == 1 + 3


def test_peer_mailbox_in_cuda_graph():
    """collect(previous) || kernels -> publish(current), replayed from one graph (what bench.py does at N > 1)."""
    from ood_object_detection_b200.distributed import PeerMailbox
    mb = PeerMailbox(DEV, local_only=True)
    src = torch.zeros(4, device=DEV)
    out, status = torch.zeros(3, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    src.copy_(torch.tensor([2.0, 1.0, 0.5, 2.0]))
    mb.publish(src)                      # prime: one record outstanding
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            mb.collect(out=out, status=status)
            src.mul_(2.0)                # stands for the step's kernels
            mb.publish(src)
    torch.cuda.current_stream().wait_stream(side)
    expect = [2.0, 1.0, 0.5, 2.0]
    for _ in range(4):
        g.replay()
        torch.cuda.synchronize()
        assert int(status) == 0
        np.testing.assert_allclose(out.cpu().numpy(), np.array(expect[:3]) / expect[3], rtol=1e-6)
        expect = [2 * v for v in expect]


def test_loss_kernel_fused_exchange_single_rank():
    """odk_loss_params.exchange: the loss kernel's finishing CTA publishes its sums and collects the previous
    step's; with world = 1 the collected value must be the ordinary (normalised) loss of that step."""
    from ood_object_detection_b200.anchors import Anchors, AnchorLabeler
    from ood_object_detection_b200 import distributed as D
    size, B, C, M = 256, 4, 20, 6
    anc = Anchors(3, 7, 3, synth.ASPECTS, 4.0, (size, size)).to(DEV)
    lab = AnchorLabeler(anc, C)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)
    unit = torch.ones((1,), device=DEV)
    mb = D.PeerMailbox(DEV, local_only=True)
    fhw = synth.feat_hw(size)
    refs = []
    for step in range(3):
        gb, gc = synth.gt_boxes(40 + step, B, size, M, C)
        co, bo = synth.head_outputs(50 + step, B, size, C, tie_free=False)
        D.local_partial_sums(lab, [torch.from_numpy(x).to(DEV) for x in co], [torch.from_numpy(x).to(DEV) for x in bo],
                             torch.from_numpy(gb).to(DEV), torch.from_numpy(gc).to(DEV), unit, mailbox=mb, **kw)
        oc, ob, onp, _, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), list(gb), list(gc))
        refs.append(orc.loss_fn(co, bo, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0))
        if step > 0:   # the launch of this step collected the sums of the previous one
            assert int(mb.status) == 0
            np.testing.assert_allclose([float(v) for v in mb.previous()], refs[step - 1], rtol=1e-5)
    got, status = mb.collect()     # drain: the last step's record
    assert int(status) == 0
    np.testing.assert_allclose([float(v) for v in got], refs[-1], rtol=1e-5)


def test_match_detections_vs_reference_goldens_and_oracle(golden):
    """odk_match_detections on a batch made of the five evaluation cases whose results the
    reference's PerImageEvaluation produced (tests/golden/evaluation.npz): per-class (score, tp/fp) sequences and
    CorLoc flags bit-identical, and per-slot labels identical to the oracle's."""
    from ood_object_detection_b200.pipeline import match_detections
    from test_oracle_golden import assert_eval_matches_golden
    g = golden('evaluation')
    for nms_iou, nms_max in ((1.0, 10000), (0.3, 50)):
        cases = [c for c in synth.EVAL_CASES if (c[5], c[6]) == (nms_iou, nms_max)]
        D, M, C = max(c[2] for c in cases), max(c[3] for c in cases), max(c[4] for c in cases)
        B = len(cases)
        dets = np.zeros((B, D, 6), np.float32)
        count = np.zeros((B,), np.int32)
        gtb = np.zeros((B, M, 4), np.float32)
        gtl = np.full((B, M), -1, np.int32)
        dif, gof = np.zeros((B, M), np.uint8), np.zeros((B, M), np.uint8)
        raw = []
        for i, (tag, seed, n_det, n_gt, Ci, _, _, flags) in enumerate(cases):
            det, scores, cls, gb, gc, d_, g_ = synth.eval_case(seed, n_det, n_gt, Ci)
            if not flags:
                d_[:] = False
                g_[:] = False
            dets[i, :n_det, 0], dets[i, :n_det, 1], dets[i, :n_det, 2], dets[i, :n_det, 3] = det[:, 1], det[:, 0], det[:, 3], det[:, 2]
            dets[i, :n_det, 4], dets[i, :n_det, 5] = scores, cls + 1            # class column is 1-based (anchors.py:157)
            count[i] = n_det
            gtb[i, :n_gt], gtl[i, :n_gt], dif[i, :n_gt], gof[i, :n_gt] = gb, gc + 1, d_, g_
            raw.append((tag, Ci, det, scores, cls, gb, gc, d_, g_))
        label, corloc = match_detections(torch.from_numpy(dets).to(DEV), torch.from_numpy(count).to(DEV), torch.from_numpy(gtb).to(DEV),
                                         torch.from_numpy(gtl).to(DEV), C, torch.from_numpy(dif).to(DEV), torch.from_numpy(gof).to(DEV),
                                         label_offset=1, nms_iou_threshold=nms_iou, nms_max_output_boxes=nms_max)
        label, corloc = label.cpu().numpy(), corloc.cpu().numpy()
        for i, (tag, Ci, det, scores, cls, gb, gc, d_, g_) in enumerate(raw):
            n = det.shape[0]
            assert (label[i, n:] == -2).all() and (corloc[i, Ci:] == 0).all()
            assert_eval_matches_golden(g, tag, Ci, scores, cls, label[i, :n], corloc[i, :Ci])
            o_label, o_corloc = orc.match_detections(det, scores, cls, gb, gc, Ci, d_, g_, 0.5, nms_iou, nms_max)
            np.testing.assert_array_equal(label[i, :n], o_label)
            np.testing.assert_array_equal(corloc[i, :Ci], o_corloc)


def test_match_detections_on_pipeline_output():
    """Detections straight from odk_postprocess (planted objects) against gt boxes placed on some of them: the
    device labels equal the oracle's for every image, without any per-image device->host hop."""
    from ood_object_detection_b200.bench import post_process_detect
    from ood_object_detection_b200.pipeline import match_detections
    size, B, C, K, D = 256, 4, 12, 1500, 40
    co, bo = synth.planted_outputs(95, B, size, C, n_obj=25)
    anc = torch.from_numpy(orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))).to(DEV)
    out = post_process_detect([torch.from_numpy(x).to(DEV) for x in co], [torch.from_numpy(x).to(DEV) for x in bo], anc, 5, C, K, D, False)
    dets, count = out['detections'], out['count']
    d = dets.cpu().numpy()
    M = 8
    gtb = np.zeros((B, M, 4), np.float32)
    gtl = np.full((B, M), -1, np.int32)
    rs = np.random.RandomState(3)
    for i in range(B):
        n = int(count[i])
        pick = rs.choice(n, size=min(M - 2, n), replace=False)
        gtb[i, :len(pick)] = d[i, pick][:, [1, 0, 3, 2]] + rs.standard_normal((len(pick), 4)).astype(np.float32) * 1.5
        gtl[i, :len(pick)] = d[i, pick, 5].astype(np.int32)
    label, corloc = match_detections(dets, count, torch.from_numpy(gtb).to(DEV), torch.from_numpy(gtl).to(DEV), C)
    assert int((label == 1).sum()) > 0
    for i in range(B):
        n = int(count[i])
        valid = gtl[i] >= 0
        o_label, o_corloc = orc.match_detections(d[i, :n][:, [1, 0, 3, 2]], d[i, :n, 4], d[i, :n, 5].astype(np.int64) - 1,
                                                 gtb[i][valid], gtl[i][valid].astype(np.int64) - 1, C)
        np.testing.assert_array_equal(label[i, :n].cpu().numpy(), o_label)
        np.testing.assert_array_equal(corloc[i].cpu().numpy(), o_corloc)

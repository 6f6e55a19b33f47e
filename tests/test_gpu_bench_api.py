"""DetBenchTrain / DetBenchPredict (reference effdet/bench.py:79-145) around a stand-in model whose
head outputs are fixed tensors: the benches must reproduce the oracle's labeler -> loss -> post-process
chain, support .backward(), and keep the reference's target-dict conventions."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class FakeModel(nn.Module):
    def __init__(self, config, cls_out, box_out):
        super().__init__()
        self.config = config
        self.cls = nn.ParameterList([nn.Parameter(torch.from_numpy(c)) for c in cls_out])
        self.box = nn.ParameterList([nn.Parameter(torch.from_numpy(b)) for b in box_out])

    def forward(self, x):
        return list(self.cls), list(self.box)


def make(size=256, B=2, C=12, soft=False, D=40, K=1500, seed=5):
    cfg = types.SimpleNamespace(num_levels=5, num_classes=C, min_level=3, max_level=7, num_scales=3,
                                aspect_ratios=synth.ASPECTS, anchor_scale=4.0, image_size=(size, size),
                                max_detection_points=K, max_det_per_image=D, soft_nms=soft, alpha=0.25, gamma=1.5,
                                delta=0.1, box_loss_weight=50.0, label_smoothing=0.0, legacy_focal=False, jit_loss=False)
    co, bo = synth.planted_outputs(seed, B, size, C, n_obj=30)
    model = FakeModel(cfg, co, bo).to(DEV)
    gb, gc = synth.gt_boxes(seed + 1, B, size, 6, C)
    return cfg, model, co, bo, gb, gc


def test_det_bench_train_and_eval():
    from ood_object_detection_b200.bench import DetBenchTrain, unwrap_bench
    size, B, C = 256, 2, 12
    cfg, model, co, bo, gb, gc = make(size, B, C)
    bench = DetBenchTrain(model).to(DEV)
    assert unwrap_bench(bench) is model
    x = torch.zeros(B, 3, size, size, device=DEV)
    target = {'bbox': torch.from_numpy(gb).to(DEV), 'cls': torch.from_numpy(gc).to(DEV),
              'img_scale': torch.ones(B, device=DEV), 'img_size': torch.full((B, 2), float(size), device=DEV)}
    bench.train()
    out = bench(x, target)
    anc = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    oc, ob, onp, _, _ = orc.batch_label_anchors(anc, list(gb), list(gc))
    fhw = synth.feat_hw(size)
    ref = orc.loss_fn(co, bo, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0, want_grad=True)
    np.testing.assert_allclose([out['loss'].item(), out['class_loss'].item(), out['box_loss'].item()], ref[:3], rtol=1e-5)
    out['loss'].backward()
    for l in range(5):
        np.testing.assert_allclose(model.cls[l].grad.cpu().numpy(), ref[3][l], rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(model.box[l].grad.cpu().numpy(), ref[4][l], rtol=2e-5, atol=1e-9)
    # eval mode adds detections; this synthetic case yields < max_det rows somewhere -> the reference's stack error
    bench.eval()
    bench.pad_detections = True
    with torch.no_grad():
        ev = bench(x, target)
    assert ev['detections'].shape == (B, cfg.max_det_per_image, 6)
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, cfg.max_detection_points)
    for i in range(B):
        det = orc.generate_detections(o_cls[i], o_box[i], anc, o_idx[i], o_klass[i], 1.0, np.array([size, size], np.float32),
                                      cfg.max_det_per_image, False)
        got = ev['detections'][i, :det.shape[0]].cpu().numpy()
        np.testing.assert_array_equal(got[:, 5], det[:, 5])
        np.testing.assert_allclose(got[:, 4], det[:, 4], rtol=1e-5)
        assert (ev['detections'][i, det.shape[0]:] == 0).all()
    # precomputed-label variant (create_labeler=False) consumes the collate's keys
    from ood_object_detection_b200.pipeline import label_batch_targets
    bench2 = DetBenchTrain(model, create_labeler=False).to(DEV).train()
    t2 = label_batch_targets(bench.anchor_labeler, dict(target), filter_valid=True)
    out2 = bench2(x, t2)
    np.testing.assert_allclose(out2['loss'].item(), out['loss'].item(), rtol=1e-6)


@pytest.mark.parametrize('soft', [False, True])
def test_det_bench_predict(soft):
    from ood_object_detection_b200.bench import DetBenchPredict
    size, B, C = 256, 2, 12
    cfg, model, co, bo, gb, gc = make(size, B, C, soft=soft)
    bench = DetBenchPredict(model).to(DEV).eval()
    bench.pad_detections = True
    x = torch.zeros(B, 3, size, size, device=DEV)
    info = {'img_scale': torch.tensor([1.0, 1.5], device=DEV), 'img_size': torch.tensor([[300., 280.], [380., 400.]], device=DEV)}
    with torch.no_grad():
        dets = bench(x, info)
        plain = bench(x)
        ood = bench.forward_with_ood(x, info)
    anc = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, cfg.max_detection_points)
    for i in range(B):
        for got_all, scale, isz in ((dets, [1.0, 1.5][i], np.array([[300., 280.], [380., 400.]], np.float32)[i]), (plain, None, None)):
            det = orc.generate_detections(o_cls[i], o_box[i], anc, o_idx[i], o_klass[i], scale, isz, cfg.max_det_per_image, soft)
            got = got_all[i, :det.shape[0]].cpu().numpy()
            np.testing.assert_array_equal(got[:, 5], det[:, 5])
            np.testing.assert_allclose(got[:, 4], det[:, 4], rtol=1e-5)
            scale_ref = np.maximum(np.abs(det[:, :4]).max(1, keepdims=True), 1.0)
            assert (np.abs(got[:, :4] - det[:, :4]) / scale_ref).max(initial=0) <= 1e-5
    assert torch.equal(ood['detections'], dets) and ood['energy'].shape == (B, cfg.max_det_per_image)
    if int(ood['count'].min()) < cfg.max_det_per_image:
        bench.pad_detections = False
        with pytest.raises(RuntimeError):
            bench(x, info)


def test_reference_script_sequence_replay_through_the_dropin_namespace():
    """The reference scripts' own call sequence, with their own import lines resolved by the drop-in `effdet`
    namespace and nothing else changed:
      preloader.py:60-62,146   Anchors / AnchorLabeler built on the HOST, batch_label_anchors on ragged lists of
                               host tensors (one image without gt) -> host targets
      pretrain.py:223-225      targets .to('cuda:0')
      pretrain.py:233-236      loss_fn(class_out, box_out, cls_anchors, bbox_anchors, num_positives); .backward()
      pretrain.py:241-249      _post_process, per-image generate_detections(...).cpu().numpy(), yxyx swap for the
                               evaluator (and the one-copy pipeline.detections_for_evaluator next to it)
      dataloader.py:210        batch_label_anchors(..., task_cls=) relabelling the caller's class tensors in place
    Every stage is compared with the CPU oracle."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(root, 'ood_object_detection_b200', 'dropin')
    saved = {k: v for k, v in sys.modules.items() if k == 'effdet' or k.startswith('effdet.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, dropin)
    try:
        anchors_mod = importlib.import_module('effdet.anchors')            # pretrain.py:14 / preloader.py:16
        loss_mod = importlib.import_module('effdet.loss')                  # pretrain.py:15
        bench_mod = importlib.import_module('effdet.bench')                # pretrain.py:13
        Anchors, AnchorLabeler, generate_detections = anchors_mod.Anchors, anchors_mod.AnchorLabeler, anchors_mod.generate_detections
        size, B, C, K = 256, 4, 15, 2000
        cfg = types.SimpleNamespace(num_levels=5, num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0,
                                    label_smoothing=0.0, legacy_focal=False, jit_loss=False, max_detection_points=K)
        anchors = Anchors(3, 7, 3, synth.ASPECTS, 4.0, (size, size))       # on the host, like the datasets build it
        labeler = AnchorLabeler(anchors, C, match_threshold=0.5)
        gb, gc = synth.gt_boxes(61, B, size, 9, C)
        lens = [9, 4, 0, 7]
        qry_bbox_ls = [torch.from_numpy(gb[i, :n].copy()) for i, n in enumerate(lens)]
        qry_cls_ls = [torch.from_numpy(gc[i, :n].copy()) for i, n in enumerate(lens)]
        cls_t, box_t, npos = labeler.batch_label_anchors(qry_bbox_ls, qry_cls_ls)      # preloader.py:146
        assert all(not t.is_cuda for t in cls_t + box_t) and not npos.is_cuda and cls_t[0].dtype == torch.int64
        anc_np = anchors.boxes.numpy()
        oc, ob, onp, _, _ = orc.batch_label_anchors(anc_np, [b.numpy() for b in qry_bbox_ls], [c.numpy() for c in qry_cls_ls])
        np.testing.assert_array_equal(torch.cat([t.reshape(B, -1) for t in cls_t], 1).numpy(), oc)
        np.testing.assert_array_equal(npos.numpy(), onp)
        qry_cls_anchors = [t.to('cuda:0') for t in cls_t]                  # pretrain.py:223-225
        qry_bbox_anchors = [t.to('cuda:0') for t in box_t]
        qry_num_positives = npos.to('cuda:0')
        co_np, bo_np = synth.planted_outputs(62, B, size, C, n_obj=25)
        class_out = [torch.from_numpy(x).to('cuda:0').requires_grad_(True) for x in co_np]
        box_out = [torch.from_numpy(x).to('cuda:0').requires_grad_(True) for x in bo_np]
        loss_fn = loss_mod.DetectionLoss(cfg)                              # pretrain.py:187
        loss, class_loss, box_loss = loss_fn(class_out, box_out, qry_cls_anchors, qry_bbox_anchors, qry_num_positives)
        loss.backward()                                                    # pretrain.py:236
        fhw = synth.feat_hw(size)
        ref = orc.loss_fn(co_np, bo_np, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0, want_grad=True)
        np.testing.assert_allclose([loss.item(), class_loss.item(), box_loss.item()], ref[:3], rtol=1e-5)
        n = float(onp.sum()) + 1.0
        g, r = class_out[0].grad.cpu().numpy(), ref[3][0]
        assert (np.abs(g - r) <= 1e-5 * np.abs(r) + 4 * 2.0 ** -24 * 0.75 / n).all()
        with torch.no_grad():                                              # pretrain.py:238-249
            class_out_post, box_out_post, indices, classes = bench_mod._post_process(
                class_out, box_out, num_levels=cfg.num_levels, num_classes=cfg.num_classes, max_detection_points=cfg.max_detection_points)
            o_cls, o_box, o_idx, o_klass = orc.post_process(co_np, bo_np, 5, C, K)
            dev_anchors = anchors.boxes.to('cuda:0')
            rows = []
            for b_ix in range(B):
                detections = generate_detections(class_out_post[b_ix], box_out_post[b_ix], dev_anchors, indices[b_ix], classes[b_ix],
                                                 None, torch.tensor([size, size]), max_det_per_image=100, soft_nms=False).cpu().numpy()
                want = orc.generate_detections(o_cls[b_ix], o_box[b_ix], anc_np, o_idx[b_ix], o_klass[b_ix], None, np.float32([size, size]), 100, False)
                from test_gpu_postprocess import assert_dets_close
                assert_dets_close(detections, want)
                rows.append(np.concatenate([detections[:, 1:2], detections[:, 0:1], detections[:, 3:4], detections[:, 2:3]], axis=1))
            # the same hand-off as one padded tensor and ONE device->host copy
            from ood_object_detection_b200.anchors import detect_batch
            from ood_object_detection_b200.pipeline import detections_for_evaluator
            dets, count, _ = detect_batch(class_out_post, box_out_post, dev_anchors, indices, classes, None, None, 100, False)
            for b_ix, rec in enumerate(detections_for_evaluator(dets, count)):
                np.testing.assert_array_equal(rec['bbox'], rows[b_ix])
        # dataloader.py:210: task_cls relabels boxes that a task-class box overlaps with IoU > 0.9 -- in the caller's tensors
        pb = [torch.tensor([[10., 10., 60., 60.], [11., 10., 60., 61.], [100., 100., 150., 160.]]), torch.tensor([[5., 5., 40., 40.]])]
        pc = [torch.tensor([3, 7, 9]), torch.tensor([7])]
        p_cls_t, _, p_npos = labeler.batch_label_anchors(pb, pc, task_cls=3)
        assert pc[0].tolist() == [3, 3, 9] and pc[1].tolist() == [7]
        oc2, _, onp2, _, _ = orc.batch_label_anchors(anc_np, [b.numpy() for b in pb], [np.array([3, 3, 9]), np.array([7])])
        np.testing.assert_array_equal(torch.cat([t.reshape(2, -1) for t in p_cls_t], 1).numpy(), oc2)
        np.testing.assert_array_equal(p_npos.numpy(), onp2)
    finally:
        sys.path.remove(dropin)
        for k in [k for k in sys.modules if k == 'effdet' or k.startswith('effdet.')]:
            del sys.modules[k]
        sys.modules.update(saved)

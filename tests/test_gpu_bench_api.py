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

"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol that
include/odk.h declares, the python shims expose the reference's API surface, argument
errors are reported through the ABI, and the drop-in `effdet` namespace resolves to our modules."""
import ctypes
import importlib
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'odk.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(odk_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from ood_object_detection_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 17
    for name in syms:
        assert hasattr(handle, name), f'libodk.so does not export {name}'
    # and the python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib().odk_version() == 3
    assert _lib.lib().odk_planar_stride(49104) == 49104 and _lib.lib().odk_planar_stride(150381) == 150384


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from ood_object_detection_b200 import _lib
    lib = _lib.lib()
    hw = _lib.int_array([16, 4])
    rc = lib.odk_assign(None, None, None, None, 2, 4, hw, 2, 9, 0.5, 1, None, None, None, 0, None)
    assert rc == -1 and b'null pointer' in lib.odk_last_error()
    rc = lib.odk_assign(None, None, None, None, 2, 4, hw, 9, 9, 0.5, 1, None, None, None, 0, None)
    assert rc == -1 and b'num_levels' in lib.odk_last_error()
    hw5 = _lib.int_array([4096, 1024, 256, 64, 16])
    assert lib.odk_topk_workspace_bytes(4, 90, hw5, 5, 9, 5000) > 4 * 16384 * 8
    assert lib.odk_postprocess_workspace_bytes(4, 90, hw5, 5, 9, 5000) > lib.odk_topk_workspace_bytes(4, 90, hw5, 5, 9, 5000)
    assert 0 < lib.odk_postprocess_flags_offset(4, 90, hw5, 5, 9, 5000) < lib.odk_postprocess_workspace_bytes(4, 90, hw5, 5, 9, 5000)
    assert lib.odk_postprocess_workspace_bytes(0, 90, hw5, 5, 9, 5000) == 0
    # odk_postprocess validates sizes before any CUDA call
    args = [None] * 26
    rc = lib.odk_postprocess(None, None, 1, 90, hw5, 5, 9, 5000, 0, None, None, None, None, 1.0, None, None, None, None, None, None,
                             None, None, None, None, None, 0, None)
    assert rc == -1 and b'null pointer' in lib.odk_last_error()
    assert lib.odk_assign_workspace_bytes(4, 100) >= 4 * 100 * 8
    with pytest.raises(RuntimeError):
        _lib.check(rc)
    # data-parallel exchange (ABI v2): sizes and argument checks
    assert lib.odk_mailbox_bytes(8) >= 2 * 8 * 32 + 8 and lib.odk_mailbox_bytes(8) % 16 == 0
    assert lib.odk_mailbox_bytes(0) == 0 and lib.odk_mailbox_bytes(_lib.MAILBOX_MAX_WORLD + 1) == 0
    assert lib.odk_partials_publish(None, None, 0, 0, None) == -1 and b'world' in lib.odk_last_error()
    assert lib.odk_partials_publish(None, None, 2, 2, None) == -1
    assert lib.odk_partials_publish(None, None, 2, 1, None) == -1 and b'partials4' in lib.odk_last_error()
    assert lib.odk_partials_collect(None, 2, None, None, 0, None) == -1
    assert lib.odk_partials_collect(None, 99, None, None, 0, None) == -1 and b'world' in lib.odk_last_error()
    # odk_detect / odk_loss validate their parameter structs before touching the device
    assert lib.odk_detect(None, None, None, None, 1, 10, None, 100, None, None, None, None, None, None, None) == -1
    assert b'params' in lib.odk_last_error()


def test_shims_refuse_cpu_tensors():
    from ood_object_detection_b200.anchors import Anchors, AnchorLabeler
    from ood_object_detection_b200.bench import _post_process
    from ood_object_detection_b200 import soft_nms as S
    anc = Anchors(3, 7, 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0, (128, 128))
    lab = AnchorLabeler(anc, 20)
    with pytest.raises(RuntimeError, match='CUDA'):
        lab.batch_label_anchors([torch.zeros(1, 4)], [torch.ones(1, dtype=torch.long)])
    with pytest.raises(RuntimeError, match='CUDA'):
        _post_process([torch.zeros(1, 9, 16, 16)], [torch.zeros(1, 36, 16, 16)], 1, 1, 10)
    with pytest.raises(RuntimeError, match='CUDA'):
        S.soft_nms(torch.zeros(3, 4), torch.zeros(3))


REFERENCE_API = {
    'anchors': ['Anchors', 'AnchorLabeler', 'decode_box_outputs', 'clip_boxes_xyxy', 'generate_detections',
                'get_feat_sizes', 'MIN_CLASS_SCORE'],
    'loss': ['loss_fn', 'DetectionLoss', 'SupportLoss', 'smooth_l1_loss', 'l2_loss', 'cosine_loss', 'huber_loss',
             'new_focal_loss', 'focal_loss_legacy', 'one_hot', 'class_loss_fn', 'box_only_loss', '_box_loss'],
    'bench': ['_post_process', '_batch_detection', 'DetBenchPredict', 'DetBenchTrain', 'unwrap_bench'],
    'soft_nms': ['soft_nms', 'batched_soft_nms', 'pairwise_iou'],
    'object_detection': ['ArgMaxMatcher', 'FasterRcnnBoxCoder', 'BoxList', 'Match', 'IouSimilarity', 'TargetAssigner'],
}


def test_reference_api_surface():
    for mod, names in REFERENCE_API.items():
        m = importlib.import_module(f'ood_object_detection_b200.{mod}')
        for n in names:
            assert hasattr(m, n), f'{mod}.{n} missing'
    from ood_object_detection_b200 import anchors, bench, loss, soft_nms
    # signatures the reference scripts rely on (pretrain.py:241-246, infer.py:683-695)
    assert list(inspect.signature(bench._post_process).parameters) == [
        'cls_outputs', 'box_outputs', 'num_levels', 'num_classes', 'max_detection_points']
    assert list(inspect.signature(anchors.generate_detections).parameters) == [
        'cls_outputs', 'box_outputs', 'anchor_boxes', 'indices', 'classes', 'img_scale', 'img_size',
        'max_det_per_image', 'soft_nms']
    assert list(inspect.signature(loss.loss_fn).parameters)[:12] == [
        'cls_outputs', 'box_outputs', 'cls_targets', 'box_targets', 'num_positives', 'num_classes', 'alpha', 'gamma',
        'delta', 'box_loss_weight', 'label_smoothing', 'legacy_focal']
    assert list(inspect.signature(anchors.AnchorLabeler.batch_label_anchors).parameters) == [
        'self', 'gt_boxes', 'gt_classes', 'filter_valid', 'task_cls']
    assert list(inspect.signature(soft_nms.soft_nms).parameters) == [
        'boxes', 'scores', 'method_gaussian', 'sigma', 'iou_threshold', 'score_threshold']
    a = anchors.Anchors(3, 7, 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0, (512, 512))
    assert a.boxes.shape == (49104, 4) and a.get_anchors_per_location() == 9 and len(a.feat_sizes) == 8
    lab = anchors.AnchorLabeler(a, 90)
    for attr in ('target_assigner', 'anchors', 'match_threshold', 'num_classes', 'indices_cache'):
        assert hasattr(lab, attr)
    assert hasattr(lab.target_assigner, '_similarity_calc')   # used by the reference at anchors.py:401


def test_dropin_namespace_resolves_to_kernels():
    """`effdet.anchors` etc. resolve to our modules when the drop-in directory is first on sys.path."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import effdet.anchors as A, effdet.loss as L, effdet.bench as B, effdet.soft_nms as S\n"
        "from effdet.object_detection import ArgMaxMatcher, FasterRcnnBoxCoder, BoxList, IouSimilarity, TargetAssigner\n"
        "from effdet.anchors import Anchors, AnchorLabeler, generate_detections\n"
        "from effdet.bench import _post_process, _batch_detection, DetBenchTrain, DetBenchPredict\n"
        "from effdet.loss import DetectionLoss, SupportLoss, smooth_l1_loss, l2_loss, cosine_loss\n"
        "import ood_object_detection_b200.anchors as OA\n"
        "assert A.Anchors is OA.Anchors and B._post_process.__module__.startswith('ood_object_detection_b200')\n"
        "print('ok')\n") % (ROOT, os.path.join(ROOT, 'ood_object_detection_b200', 'dropin'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == 'ok', out.stderr[-2000:]


def test_matcher_and_coder_tensor_paths():
    """The non-hot-path tensor implementations agree with the oracle on CPU tensors."""
    import numpy as np
    import synth
    from oracle import oracle as orc
    from ood_object_detection_b200.object_detection import ArgMaxMatcher, BoxList, FasterRcnnBoxCoder, Match
    anc = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (128, 128))
    gb, gc = synth.gt_boxes(3, 1, 128, 7, 20)
    sim = torch.from_numpy(orc.iou_matrix(gb[0], anc))
    m = ArgMaxMatcher(0.5, 0.5, True, True).match(sim)
    _, _, _, om, _ = orc.batch_label_anchors(anc, [gb[0]], [gc[0]])
    np.testing.assert_array_equal(m.match_results.numpy(), om[0])
    assert isinstance(m, Match) and m.num_matched_columns() == int((om[0] >= 0).sum())
    assert ArgMaxMatcher(0.5, 0.5, True, True).match(sim[:0]).match_results.eq(-1).all()
    pos = np.nonzero(om[0] >= 0)[0]
    enc = FasterRcnnBoxCoder().encode(BoxList(torch.from_numpy(gb[0][om[0][pos]])), BoxList(torch.from_numpy(anc[pos])))
    _, ob, _, _, _ = orc.batch_label_anchors(anc, [gb[0]], [gc[0]])
    np.testing.assert_allclose(enc.numpy(), ob[0][pos], rtol=1e-5, atol=1e-7)
    dec = FasterRcnnBoxCoder().decode(enc, BoxList(torch.from_numpy(anc[pos]))).boxes()
    np.testing.assert_allclose(dec.numpy(), gb[0][om[0][pos]], rtol=1e-4, atol=1e-3)
    with pytest.raises(ValueError):
        ArgMaxMatcher(0.4, 0.5)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors of the by-pointer parameter structs must have the C compiler's layout."""
    import ctypes
    import subprocess
    from ood_object_detection_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "odk.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(odk_loss_params), '
                   'offsetof(odk_loss_params, exchange), sizeof(odk_exchange), offsetof(odk_exchange, world), '
                   'offsetof(odk_exchange, num_pos_plus_1), offsetof(odk_exchange, status), sizeof(odk_detect_params), '
                   'offsetof(odk_loss_params, clear_keys), offsetof(odk_exchange, normalized), '
                   'offsetof(odk_exchange, timeout_ms), offsetof(odk_detect_params, pipeline)); return 0; }\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', str(exe), str(src)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.LossParams), _lib.LossParams.exchange.offset, ctypes.sizeof(_lib.Exchange),
            _lib.Exchange.world.offset, _lib.Exchange.num_pos_plus_1.offset, _lib.Exchange.status.offset,
            ctypes.sizeof(_lib.DetectParams), _lib.LossParams.clear_keys.offset, _lib.Exchange.normalized.offset,
            _lib.Exchange.timeout_ms.offset, _lib.DetectParams.pipeline.offset]
    assert got == want


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) must print ONE JSON line with
    the contract's keys; it runs on the host only, so it is checked here without a GPU."""
    import json
    import subprocess
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                                   '--warmup', '1'], cwd=ROOT, timeout=600).decode()
    lines = [l for l in out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'images/s'
    for key in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data',
                'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['e2e']['value'] == d['value'] and d['value'] > 0
    assert 'workload' in d['config']


def test_bench_wrappers_with_stubbed_kernels(monkeypatch):
    """DetBenchPredict / DetBenchTrain host logic (reference bench.py:79-145) with the kernel entry points
    stubbed out: which tensors reach which call, the output dict, the label_* target path, eval-mode
    detections, unwrap_bench.  (The real kernels behind them are covered by the GPU tests.)"""
    import types
    from ood_object_detection_b200 import bench as B
    cfg = types.SimpleNamespace(min_level=3, max_level=7, num_levels=5, num_scales=3, aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)],
                                anchor_scale=4.0, image_size=(128, 128), num_classes=7, max_detection_points=50,
                                max_det_per_image=5, soft_nms=True, alpha=0.25, gamma=1.5, delta=0.1,
                                box_loss_weight=50.0, label_smoothing=0.0, legacy_focal=False, jit_loss=False)

    class Model(torch.nn.Module):
        config = cfg

        def forward(self, x):
            return ['cls_levels', x.shape[0]], ['box_levels']

    calls = {}

    def fake_chain(cls_outputs, box_outputs, anchor_boxes, num_levels, num_classes, max_detection_points=5000,
                   max_det_per_image=100, soft_nms=False, img_scale=None, img_size=None, **kw):
        calls['det'] = (cls_outputs, box_outputs, tuple(anchor_boxes.shape), num_levels, num_classes, max_detection_points,
                        max_det_per_image, soft_nms, img_scale, img_size)
        n = cls_outputs[1]
        return {'detections': torch.full((n, max_det_per_image, 6), 2.0),
                'count': torch.full((n,), calls.get('count', max_det_per_image), dtype=torch.int32)}

    monkeypatch.setattr(B, 'post_process_detect', fake_chain)
    x = torch.zeros(3, 3, 128, 128)

    pred = B.DetBenchPredict(Model())
    assert (pred.num_levels, pred.num_classes, pred.max_detection_points, pred.max_det_per_image, pred.soft_nms) == (5, 7, 50, 5, True)
    assert pred.config is cfg and tuple(pred.anchors.boxes.shape) == (3069, 4)
    assert tuple(pred(x).shape) == (3, 5, 6)
    assert calls['det'] == (['cls_levels', 3], ['box_levels'], (3069, 4), 5, 7, 50, 5, True, None, None)
    calls['count'] = 4          # an image with fewer than max_det rows: the reference's torch.stack raises (bench.py:76)
    with pytest.raises(RuntimeError, match='equal size'):
        pred(x)
    pred.pad_detections = True
    assert tuple(pred(x, {'img_scale': 'scale', 'img_size': 'size'}).shape) == (3, 5, 6)
    assert calls['det'][8:] == ('scale', 'size')
    calls['count'] = 5

    monkeypatch.setattr(B, 'detect_with_ood', lambda *a, **k: ('ood', a[0], a[1], tuple(a[2].shape), a[3:], k))
    assert pred.forward_with_ood(x, {'img_scale': 'scale', 'img_size': 'size'}, temperature=2.0) == \
        ('ood', ['cls_levels', 3], ['box_levels'], (3069, 4), (5, 7, 50, 5, True, 'scale', 'size', 2.0), {})

    train = B.DetBenchTrain(Model(), create_labeler=False)
    assert train.anchor_labeler is None
    seen = {}

    class FakeLoss(torch.nn.Module):
        def forward(self, cls_out, box_out, cls_t, box_t, npos):
            seen['loss'] = (cls_out, box_out, cls_t, box_t, npos)
            return 'total', 'cls', 'box'

        def forward_fused(self, cls_out, box_out, label_batch):
            seen['fused'] = (cls_out, box_out, label_batch)
            return 't', 'c', 'b'

    train.loss_fn = FakeLoss()
    target = {f'label_cls_{l}': f'c{l}' for l in range(5)}
    target.update({f'label_bbox_{l}': f'b{l}' for l in range(5)})
    target.update(label_num_positives='npos', img_scale='scale', img_size='size')
    train.train()
    out = train(x, target)
    assert out == {'loss': 'total', 'class_loss': 'cls', 'box_loss': 'box'}
    assert seen['loss'] == (['cls_levels', 3], ['box_levels'], [f'c{l}' for l in range(5)], [f'b{l}' for l in range(5)], 'npos')
    train.eval()
    out = train(x, target)
    assert tuple(out['detections'].shape) == (3, 5, 6) and calls['det'][8:] == ('scale', 'size')
    with pytest.raises(AssertionError):
        train(x, {'bbox': None, 'cls': None})

    with_labeler = B.DetBenchTrain(Model())   # default: own labeler, fused path
    assert with_labeler.anchor_labeler is not None and with_labeler.anchor_labeler.num_classes == 7
    monkeypatch.setattr(with_labeler.anchor_labeler, 'assign', lambda boxes, cls, transient=False: ('label_batch', boxes, cls))
    with_labeler.loss_fn = FakeLoss()
    with_labeler.train()
    assert with_labeler(x, {'bbox': 'gtb', 'cls': 'gtc'}) == {'loss': 't', 'class_loss': 'c', 'box_loss': 'b'}
    assert seen['fused'] == (['cls_levels', 3], ['box_levels'], ('label_batch', 'gtb', 'gtc'))

    wrapped = types.SimpleNamespace(module=train)
    assert isinstance(B.unwrap_bench(wrapped), Model)

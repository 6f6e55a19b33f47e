"""Times every BASELINE.json configuration on one GPU (CUDA events, inputs resident in HBM) and
prints one JSON line per configuration.  Not the headline bench (bench.py): supporting numbers for
profiles/ and DESIGN.md."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from ood_object_detection_b200.anchors import Anchors, AnchorLabeler, detect_batch  # noqa: E402
from ood_object_detection_b200.bench import _post_process, detect_with_ood  # noqa: E402
from ood_object_detection_b200.loss import loss_fn_fused  # noqa: E402

DEV = torch.device('cuda:0')
PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
KW = dict(alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)


def outputs(seed, B, size, C):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    feat = synth.feat_hw(size)
    return ([torch.randn((B, 9 * C, h, w), generator=g, device=DEV) * 1.5 - 4.6 for h, w in feat],
            [torch.randn((B, 36, h, w), generator=g, device=DEV) * 0.2 for h, w in feat])


def timeit(fn, iters=20, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def train_config(tag, name, B, M, C=90, grad=False):
    size, scale = synth.MODEL_SHAPES[name]
    anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(DEV)
    lab = AnchorLabeler(anc, C)
    gb, gc = synth.gt_boxes(1, B, size, M, C)
    gb, gc = torch.from_numpy(gb).to(DEV), torch.from_numpy(gc).to(DEV)
    cls, box = outputs(1, B, size, C)
    A = anc.boxes.shape[0]
    if grad:
        cls = [c.requires_grad_(True) for c in cls]
        box = [b.requires_grad_(True) for b in box]

    def step():
        lb = lab.assign(gb, gc)
        tot, _, _ = loss_fn_fused(cls, box, lb, num_classes=C, **KW)
        if grad:
            tot.backward()
            for t in cls + box:
                t.grad = None
    with torch.set_grad_enabled(grad):
        ms = timeit(step, iters=10 if B * A > 2e7 else 20)
        t_assign = timeit(lambda: lab.assign(gb, gc), iters=10)
    by = B * A * (4 * C + 16) * (2 if grad else 1)
    print(json.dumps({'config': tag, 'workload': f'{name} B={B} M={M} C={C} labeler+fused loss {"fwd+grad" if grad else "fwd"}',
                      'ms_per_step': ms, 'images_per_s': B / ms * 1e3, 'assign_ms': t_assign,
                      'GBps_whole_step': by / ms / 1e6, 'frac_of_measured_peak_whole_step': by / ms / 1e6 / PEAK}))


def pp_config(tag, name, B, soft, ood=False, C=90, K=5000, D=100):
    size, scale = synth.MODEL_SHAPES[name]
    anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(DEV)
    cls, box = outputs(2, B, size, C)
    A = anc.boxes.shape[0]

    def step():
        if ood:
            return detect_with_ood(cls, box, anc.boxes, 5, C, K, D, soft)
        pp = _post_process(cls, box, 5, C, K)
        return detect_batch(pp[0], pp[1], anc.boxes, pp[2], pp[3], None, None, D, soft)
    ms = timeit(step, iters=10)
    t_topk = timeit(lambda: _post_process(cls, box, 5, C, K), iters=10)
    by = B * A * 4 * C
    print(json.dumps({'config': tag, 'workload': f'{name} B={B} C={C} top-{K} + decode + {"soft-" if soft else ""}NMS-{D}{" + OOD" if ood else ""}',
                      'ms_per_step': ms, 'images_per_s': B / ms * 1e3, 'topk_ms': t_topk,
                      'GBps_whole_step': by / ms / 1e6, 'frac_of_measured_peak_whole_step': by / ms / 1e6 / PEAK}))


if __name__ == '__main__':
    with torch.no_grad():
        train_config('configs[1]', 'd0', 64, 10)
    train_config('configs[1]+grad', 'd0', 64, 10, grad=True)
    with torch.no_grad():
        train_config('configs[0]-shape', 'd0', 8, 10)
        pp_config('configs[0]-shape', 'd0', 8, False)
        pp_config('configs[2]', 'd3', 32, False)
        pp_config('configs[2]', 'd3', 32, True)
        pp_config('configs[3]', 'd5', 32, False, ood=True)
        train_config('configs[4]', 'd7', 128, 100)
    torch.cuda.empty_cache()
    train_config('configs[4]+grad', 'd7', 64, 100, grad=True)

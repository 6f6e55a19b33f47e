"""Labeler alone at BASELINE.json configs[4] (D7 1536^2, B=128, 100 gt/img) and configs[1] (D0 B=64 M=10): device time per
assign() call, anchors recomputed from the float64 generator vs gathered from the table."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from ood_object_detection_b200.anchors import Anchors, AnchorLabeler  # noqa: E402

dev = torch.device('cuda:0')
for name, B, M in (('d0', 64, 10), ('d7', 128, 100)):
    size, scale = synth.MODEL_SHAPES[name]
    labeler = AnchorLabeler(Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev), 90, match_threshold=0.5)
    gb, gc = synth.gt_boxes(100, B, size, M, 90)
    gb, gc = torch.from_numpy(gb).to(dev), torch.from_numpy(gc).to(dev)
    for gen in (True, False):
        labeler.use_anchor_generator = gen
        for _ in range(3):
            lb = labeler.assign(gb, gc)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            lb = labeler.assign(gb, gc)
        b.record()
        torch.cuda.synchronize()
        print(f'{name} B={B} M={M} anchor generator={gen}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per assign() (memset + kernel), '
              f'positives {int(lb.num_positives.sum())}')

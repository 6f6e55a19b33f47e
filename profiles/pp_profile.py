"""Small driver for profiling the post-process path (BASELINE.json configs[2]: D3 896^2, B=32):
3 iterations of odk_topk + odk_detect on synthetic head outputs.  Run plain, then under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from ood_object_detection_b200.anchors import Anchors, detect_batch  # noqa: E402
from ood_object_detection_b200.bench import _post_process  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'd3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
size, scale = synth.MODEL_SHAPES[name]
C, K, D = 90, 5000, 100
dev = torch.device('cuda:0')
g = torch.Generator(device=dev)
g.manual_seed(3)
feat = synth.feat_hw(size)
cls_out = [torch.randn((B, 9 * C, h, w), generator=g, device=dev) * 1.5 - 4.6 for h, w in feat]
box_out = [torch.randn((B, 36, h, w), generator=g, device=dev) * 0.2 for h, w in feat]
anchors = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev)
for soft in (False, True):
    for it in range(iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        pp = _post_process(cls_out, box_out, 5, C, K)
        e1.record()
        dets, count, src = detect_batch(pp[0], pp[1], anchors.boxes, pp[2], pp[3], None, None, D, soft)
        e2.record()
        torch.cuda.synchronize()
        print(f'{name} B={B} soft={soft} iter {it}: topk {e0.elapsed_time(e1):.3f} ms, detect {e1.elapsed_time(e2):.3f} ms, '
              f'kept min {int(count.min())}')

"""Small driver for profiling the training-target path (BASELINE.json configs[1]: D0 512^2, B=64, C=90,
M=10): 3 iterations of labeler + fused loss (forward), then 3 of forward + gradient.  Run plain, then
under ncu.  Second argument `cl`: the head outputs are channels_last (flat stream + patch kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from ood_object_detection_b200.anchors import Anchors, AnchorLabeler  # noqa: E402
from ood_object_detection_b200.loss import loss_fn_fused  # noqa: E402

size, scale = synth.MODEL_SHAPES['d0']
B, C, M = 64, 90, 10
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda:0')
g = torch.Generator(device=dev)
g.manual_seed(1)
feat = synth.feat_hw(size)
cls_out = [torch.randn((B, 9 * C, h, w), generator=g, device=dev) * 1.5 - 4.6 for h, w in feat]
box_out = [torch.randn((B, 36, h, w), generator=g, device=dev) * 0.2 for h, w in feat]
if 'cl' in sys.argv[2:]:
    cls_out = [t.contiguous(memory_format=torch.channels_last) for t in cls_out]
    box_out = [t.contiguous(memory_format=torch.channels_last) for t in box_out]
gb, gc = synth.gt_boxes(100, B, size, M, C)
gb, gc = torch.from_numpy(gb).to(dev), torch.from_numpy(gc).to(dev)
labeler = AnchorLabeler(Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev), C, match_threshold=0.5)
TRANSIENT = 'transient' in sys.argv[2:]   # the bench's mode: the loss walks the labeler's list and clears the keys
labeler.use_anchor_generator = os.environ.get('ODK_TEST_ANCHOR_GEN', '1') == '1'
kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)
for grad in (False, True):
    for t in cls_out + box_out:
        t.requires_grad_(grad)
    for it in range(iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        lb = labeler.assign(gb, gc, transient=TRANSIENT)
        e1.record()
        with torch.set_grad_enabled(grad):
            tot, cl, bx = loss_fn_fused(cls_out, box_out, lb, **kw)
        if grad:
            tot.backward()
            for t in cls_out + box_out:
                t.grad = None
        e2.record()
        torch.cuda.synchronize()
        print(f'd0 B={B} grad={grad} iter {it}: assign {e0.elapsed_time(e1):.3f} ms, loss {e1.elapsed_time(e2):.3f} ms, '
              f'total {float(tot.detach()):.4f}')

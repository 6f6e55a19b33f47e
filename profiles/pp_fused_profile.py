"""Driver for the pipelined post-process (odk_postprocess) on BASELINE.json configs[2] (D3 896^2, B=32) and
configs[3] (D5 1280^2, B=32, + OOD): the fused entry point next to the separate odk_topk + odk_detect chain,
dense (i.i.d.) and planted (trained-net-like) score regimes, hard and soft NMS.  Prints medians and the
per-image flag array (how many images left the sampled path).  Run plain, then under ncu.

    python profiles/pp_fused_profile.py [d3|d5] [B] [iters] [regimes: dense,planted]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from ood_object_detection_b200.anchors import Anchors, detect_batch  # noqa: E402
from ood_object_detection_b200.bench import _post_process, post_process_detect  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'd3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
regimes = (sys.argv[4] if len(sys.argv) > 4 else 'dense,planted').split(',')
pipelines = (sys.argv[5] if len(sys.argv) > 5 else 'staged,persistent').split(',')
show_timeline = os.environ.get('ODK_SHOW_TIMELINE', '1') == '1'
size, scale = synth.MODEL_SHAPES[name]
C, K, D = 90, 5000, 100
dev = torch.device('cuda:0')
feat = synth.feat_hw(size)
anchors = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev)
A = anchors.boxes.shape[0]
GB = B * A * C * 4 / 1e9


def make(regime):
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    if regime == 'dense':
        cls = [torch.randn((B, 9 * C, h, w), generator=g, device=dev) * 1.5 - 4.6 for h, w in feat]
    else:   # background N(-7, 1), 50 objects x 20 anchors at +3 per image (synth.planted_outputs on device)
        cls = [torch.randn((B, 9 * C, h, w), generator=g, device=dev) - 7.0 for h, w in feat]
        rs = np.random.RandomState(11)
        for i in range(B):
            for _ in range(50):
                lvl = rs.randint(0, 3)
                h, w = feat[lvl]
                y0, x0, c = rs.randint(0, h), rs.randint(0, w), rs.randint(0, C)
                for _ in range(20):
                    a = rs.randint(0, 9)
                    y = min(h - 1, max(0, y0 + rs.randint(-1, 2)))
                    x = min(w - 1, max(0, x0 + rs.randint(-1, 2)))
                    cls[lvl][i, a * C + c, y, x] = 3.0 + rs.standard_normal() * 0.5
    box = [torch.randn((B, 36, h, w), generator=g, device=dev) * 0.2 for h, w in feat]
    return cls, box


def timed(fn):
    """(CUDA-graph replay ms per call, back-to-back eager ms per call, worst single eager call ms, last result)."""
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    eager = a.elapsed_time(b) / iters
    worst = 0.0
    for _ in range(iters):
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        worst = max(worst, a.elapsed_time(b))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            gout = fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, eager, worst, gout


for regime in regimes:
    cls_out, box_out = make(regime)
    for soft in ((True,) if os.environ.get('ODK_PP_SOFT') == '1' else (False, True)):
        for ood in ((False, True) if name == 'd5' else (False,)):
            def chain():
                pp = _post_process(cls_out, box_out, 5, C, K)
                return detect_batch(pp[0], pp[1], anchors.boxes, pp[2], pp[3], None, None, D, soft)
            mc, ec, wc, ref = timed(chain)
            print(f'{name} B={B} {regime} soft={soft} ood={ood}: odk_topk + odk_detect chain: graph {mc:.4f} ms (eager {ec:.4f}, worst call {wc:.4f})', flush=True)
            for pipe in pipelines:
                def fused():
                    return post_process_detect(cls_out, box_out, anchors.boxes, 5, C, K, D, soft, with_ood=ood, return_flags=True,
                                               pipeline=pipe)
                mf, ef, wf, out = timed(fused)
                same = bool(torch.equal(out['detections'], ref[0]) and torch.equal(out['count'], ref[1]))
                print(f'  odk_postprocess[{pipe}]: graph {mf:.4f} ms (eager {ef:.4f}, worst call {wf:.4f}) = {GB / mf * 1e3:.0f} GB/s '
                      f'= {GB / mf * 1e3 / 6521.4:.3f} of peak | equal to chain {same} | kept min {int(out["count"].min())} | '
                      f'flagged images {int(out["flags"].sum())}/{B}', flush=True)
                if not show_timeline:
                    continue
                tl = out['timeline'].cpu().numpy().astype(np.int64)
                ev = tl[:-2].reshape(B, 16).astype(np.float64)
                t0 = float(tl[-2]) if pipe == 'persistent' else float(ev[:, 2].min())
                ev = (ev - t0) / 1e3
                if pipe == 'persistent':
                    print(f'    persistent kernel (start -> last CTA out): {(tl[-1] - tl[-2]) / 1e3:.1f} us')
                print('    per image, us: streamed claimed | tail start..end = select (hist scans scatter rank rest) filter decode offsets nms rows')
                raw = tl[:-2].reshape(B, 16)
                print('    candidates per image:', raw[:, 13].tolist())
                print('    survivors (ranked):  ', [(int(v & 0xFFFFFFFF), int(v >> 32)) for v in raw[:, 14]])
                order = np.argsort(ev[:, 2])
                for i in list(order[:3]) + list(order[-3:]):
                    e = ev[i]
                    print(f'    {i:2d}: {e[0]:7.1f} {e[1]:7.1f} | {e[2]:7.1f}..{e[3]:7.1f} = {e[3] - e[2]:6.1f}: select {e[4] - e[2]:5.1f} '
                          f'({e[9] - e[2]:4.1f} {e[10] - e[9]:4.1f} {e[11] - e[10]:4.1f} {e[12] - e[11]:4.1f} {e[4] - e[12]:4.1f}) '
                          f'filter {e[5] - e[4]:5.1f} decode {e[6] - e[5]:5.1f} offsets {e[7] - e[6]:5.1f} nms {e[8] - e[7]:5.1f} rows {e[3] - e[8]:5.1f}')

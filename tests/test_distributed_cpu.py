"""N>1 host logic on CPU: world_size-2 gloo processes run the sharding + collectives of
ood_object_detection_b200.distributed with the CPU oracle standing in for the kernels, and must
reproduce the single-process result at the global batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import synth
    from oracle import oracle as orc
    from ood_object_detection_b200 import distributed as D
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    size, B, C, M = 128, 6, 12, 5
    anchors = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    gb, gc = synth.gt_boxes(5, B, size, M, C)
    co, bo = synth.head_outputs(6, B, size, C, tie_free=False)
    fhw = synth.feat_hw(size)
    lo, hi = D.shard_range(B, rank, world)
    # ---- loss: local shard against the GLOBAL normaliser, then all-reduce of the partial sums ----
    cls_t, box_t, npos, _, _ = orc.batch_label_anchors(anchors, list(gb[lo:hi]), list(gc[lo:hi]))
    norm = D.global_normalizer(torch.from_numpy(npos))
    # the oracle takes num_positives and adds 1 itself: pass the global count
    fake_npos = np.array([norm.item() - 1.0], np.float32)
    part = orc.loss_fn([c[lo:hi] for c in co], [b[lo:hi] for b in bo], orc.split_levels(cls_t, fhw),
                       orc.split_levels(box_t, fhw), fake_npos, C, 0.25, 1.5, 0.1, 50.0)
    tot, cl, bl = D.reduce_losses(*[torch.tensor(v, dtype=torch.float64) for v in part])
    # forward-only form: unit normaliser locally, one collective, divide by the global N afterwards
    unit = orc.loss_fn([c[lo:hi] for c in co], [b[lo:hi] for b in bo], orc.split_levels(cls_t, fhw),
                       orc.split_levels(box_t, fhw), np.zeros(1, np.float32), C, 0.25, 1.5, 0.1, 50.0)
    one = D.forward_losses_one_collective(torch.tensor(unit[1]), torch.tensor(unit[2]), torch.from_numpy(npos), 50.0)
    # packed form: [cls + w*box, cls, box, sum(num_positives) + 1] as the two kernels leave it in one buffer
    buf = torch.tensor([unit[1] + 50.0 * unit[2], unit[1], unit[2], float(npos.sum()) + 1.0], dtype=torch.float32)
    packed = D.all_reduce_partial_sums(buf, async_op=True).result()
    # ---- detections: each rank post-processes its images, all-gather in rank order ----
    o_cls, o_box, o_idx, o_klass = orc.post_process([c[lo:hi] for c in co], [b[lo:hi] for b in bo], 5, C, 300)
    Dmax = 20
    dets = torch.zeros((hi - lo, Dmax, 6))
    count = torch.zeros((hi - lo,), dtype=torch.int32)
    for i in range(hi - lo):
        d = orc.generate_detections(o_cls[i], o_box[i], anchors, o_idx[i], o_klass[i], None, None, Dmax, False)
        dets[i, :d.shape[0]] = torch.from_numpy(d)
        count[i] = d.shape[0]
    g_dets, g_count = D.gather_detections(dets, count)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'sharded.npz'), loss=np.array([tot.item(), cl.item(), bl.item()]),
                 one=np.array([float(v) for v in one]), packed=np.array([float(v) for v in packed]),
                 dets=g_dets.numpy(), count=g_count.numpy(), norm=norm.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_path_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), 'sharded.npz'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import synth
    from oracle import oracle as orc
    size, B, C, M = 128, 6, 12, 5
    anchors = orc.anchor_boxes(3, 7, 3, synth.ASPECTS, 4.0, (size, size))
    gb, gc = synth.gt_boxes(5, B, size, M, C)
    co, bo = synth.head_outputs(6, B, size, C, tie_free=False)
    fhw = synth.feat_hw(size)
    cls_t, box_t, npos, _, _ = orc.batch_label_anchors(anchors, list(gb), list(gc))
    ref = orc.loss_fn(co, bo, orc.split_levels(cls_t, fhw), orc.split_levels(box_t, fhw), npos, C, 0.25, 1.5, 0.1, 50.0)
    np.testing.assert_allclose(got['norm'], npos.sum() + 1.0)
    np.testing.assert_allclose(got['loss'], ref, rtol=1e-6)   # same element values, different summation split
    np.testing.assert_allclose(got['one'], ref, rtol=1e-5)    # one-collective forward form (fp32 wire format)
    np.testing.assert_allclose(got['packed'], ref, rtol=1e-5)  # same, 4-float buffer written by the kernels
    o_cls, o_box, o_idx, o_klass = orc.post_process(co, bo, 5, C, 300)
    assert got['dets'].shape == (B, 20, 6)
    for i in range(B):
        d = orc.generate_detections(o_cls[i], o_box[i], anchors, o_idx[i], o_klass[i], None, None, 20, False)
        assert got['count'][i] == d.shape[0]
        np.testing.assert_array_equal(got['dets'][i, :d.shape[0]], d)


def test_shard_range_partitions_every_batch():
    from ood_object_detection_b200.distributed import shard_range
    for B in (1, 7, 8, 32, 129):
        for W in (1, 2, 4, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_helpers_are_identity():
    from ood_object_detection_b200 import distributed as D
    npos = torch.tensor([3.0, 0.0, 5.0])
    assert D.global_normalizer(npos).item() == 9.0
    t = D.reduce_losses(torch.tensor(1.5), torch.tensor(1.0), torch.tensor(0.01))
    assert [float(x) for x in t] == [1.5, 1.0, pytest.approx(0.01)]
    tot, cl, bx = D.all_reduce_partial_sums(torch.tensor([12.0, 2.0, 0.2, 4.0]))
    assert [float(x) for x in (tot, cl, bx)] == [3.0, 0.5, pytest.approx(0.05)]
    d, c = D.gather_detections(torch.zeros(2, 4, 6), torch.zeros(2, dtype=torch.int32))
    assert d.shape == (2, 4, 6) and c.shape == (2,)

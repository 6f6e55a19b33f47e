"""GPU parity: odk_assign / odk_targets / odk_loss (through the effdet-API shims and the C ABI)
against the CPU oracle and the reference goldens.  Bars: assignments / class targets /
num_positives bit-exact; encoded boxes, losses and gradients within 1e-5 relative (fp32)."""
import numpy as np
import pytest
import torch

import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def dev():
    return torch.device('cuda:0')


USE_GRID = [True]   # flipped by the `kernel` fixture: every labeler test runs on both assignment kernels


@pytest.fixture(autouse=True, params=['grid', 'dense'])
def kernel(request):
    USE_GRID[0] = request.param == 'grid'
    yield request.param


def make_labeler(size, scale=4.0, num_classes=90, thr=0.5):
    from ood_object_detection_b200.anchors import Anchors, AnchorLabeler
    anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev())
    lab = AnchorLabeler(anc, num_classes, match_threshold=thr)
    lab.use_grid_kernel = USE_GRID[0]
    return anc, lab


def flat_targets(cls_t, box_t):
    B = cls_t[0].shape[0]
    c = torch.cat([t.reshape(B, -1) for t in cls_t], 1).cpu().numpy()
    b = torch.cat([t.reshape(B, -1, 4) for t in box_t], 1).cpu().numpy()
    return c, b


def check_against_oracle(anc, lab, gb_list, gc_list, **kw):
    gbt = [torch.from_numpy(np.asarray(b, np.float32).reshape(-1, 4)).to(dev()) for b in gb_list]
    gct = [torch.from_numpy(np.asarray(c)).to(dev()) for c in gc_list]
    cls_t, box_t, npos = lab.batch_label_anchors(gbt, gct, **kw)
    assert cls_t[0].dtype == torch.int64 and box_t[0].dtype == torch.float32
    c, b = flat_targets(cls_t, box_t)
    oc, ob, onp, _, ocls = orc.batch_label_anchors(anc.boxes.cpu().numpy(), gb_list, gc_list,
                                                    match_threshold=lab.match_threshold, **kw)
    np.testing.assert_array_equal(c, oc)
    np.testing.assert_array_equal(npos.cpu().numpy(), onp)
    assert ((b != 0) == (ob != 0)).all()
    np.testing.assert_allclose(b, ob, rtol=RTOL, atol=1e-7)
    return c, b, npos, gct


@pytest.mark.parametrize('tag,kw', [('empty', {}), ('zero_iou', {}), ('identical', {}), ('tiny', {}),
                                    ('padded_float', {}), ('collide', {}), ('nofilter', {'filter_valid': False})])
def test_labeler_kat(golden, tag, kw):
    g = golden('labeler')
    anc, lab = make_labeler(512)
    c, b, npos, _ = check_against_oracle(anc, lab, [g[f'kat_{tag}_boxes']], [g[f'kat_{tag}_classes']], **kw)
    pos = np.nonzero(c[0] != -1)[0]
    np.testing.assert_array_equal(pos, g[f'kat_{tag}_pos_idx'])
    np.testing.assert_array_equal(c[0][pos], g[f'kat_{tag}_pos_cls'])
    np.testing.assert_allclose(b[0][pos], g[f'kat_{tag}_pos_box'], rtol=RTOL, atol=1e-7)
    np.testing.assert_array_equal(npos.cpu().numpy(), g[f'kat_{tag}_npos'])


@pytest.mark.parametrize('tag,size,m,seed,integer', [('r256_m10', 256, 10, 11, False), ('r256_m100', 256, 100, 12, False),
                                                     ('r256_int', 256, 24, 13, True), ('r512_m10', 512, 10, 14, False)])
def test_labeler_golden_dense_tensor_input(golden, tag, size, m, seed, integer):
    """[B, M, 4] / [B, M] tensor input with -1 padded classes (the collate format)."""
    g = golden('labeler')
    anc, lab = make_labeler(size)
    gb, _ = synth.gt_boxes(seed, 3, size, m, 90, integer=integer)
    gc = g[f'{tag}_gc']
    cls_t, box_t, npos = lab.batch_label_anchors(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()))
    c, b = flat_targets(cls_t, box_t)
    np.testing.assert_array_equal(c, g[f'{tag}_cls'].astype(np.int64))
    np.testing.assert_array_equal(npos.cpu().numpy(), g[f'{tag}_npos'])
    np.testing.assert_allclose(b, g[f'{tag}_box'], rtol=RTOL, atol=1e-7)
    assert ((b != 0) == (g[f'{tag}_box'] != 0)).all()


@pytest.mark.parametrize('thr', [0.4, 0.7])
def test_labeler_thresholds(golden, thr):
    g = golden('labeler')
    anc, lab = make_labeler(256, thr=thr)
    gb, gc = synth.gt_boxes(15, 2, 256, 20, 90)
    c, b, npos, _ = check_against_oracle(anc, lab, list(gb), list(gc))
    np.testing.assert_array_equal(c, g[f'thr{int(thr * 10)}_cls'].astype(np.int64))


def test_labeler_task_cls(golden):
    g = golden('labeler')
    anc, lab = make_labeler(256)
    c, b, npos, gct = check_against_oracle(anc, lab, list(g['task_boxes']), list(g['task_classes_in']), task_cls=5)
    np.testing.assert_array_equal(gct[0].cpu().numpy(), g['task_classes_out'][0])  # mutated in place like the reference
    np.testing.assert_array_equal(c, g['task_cls'].astype(np.int64))


@pytest.mark.parametrize('name,B,m', [('d0', 4, 10), ('d0', 2, 100), ('d3', 2, 37), ('d7', 2, 100)])
def test_labeler_model_shapes(name, B, m):
    """Full-size anchor tables, ragged lists incl. an empty image and degenerate boxes."""
    size, scale = synth.MODEL_SHAPES[name]
    anc, lab = make_labeler(size, scale)
    gb, gc = synth.gt_boxes(900 + B + m, B, size, m, 90)
    gbl, gcl = [gb[i] for i in range(B)], [gc[i] for i in range(B)]
    gbl[0], gcl[0] = gbl[0][:0], gcl[0][:0]                      # no gt at all
    gbl[1] = gbl[1].copy()
    gbl[1][0] = [10, 10, 10, 10]                                 # zero-area box -> forced onto anchor 0
    gbl[1][1] = [size * 4, size * 4, size * 5, size * 5]         # outside the image
    check_against_oracle(anc, lab, gbl, gcl)


def test_target_assigner_api():
    """effdet.object_detection API: arbitrary anchor BoxList through TargetAssigner.assign."""
    from ood_object_detection_b200.object_detection import BoxList
    anc, lab = make_labeler(256)
    gb, gc = synth.gt_boxes(77, 1, 256, 9, 90)
    cls_t, reg_t, match = lab.target_assigner.assign(BoxList(anc.boxes), BoxList(torch.from_numpy(gb[0]).to(dev())),
                                                     torch.from_numpy(gc[0]).to(dev()))
    oc, ob, onp, om, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), [gb[0]], [gc[0]])
    np.testing.assert_array_equal(match.match_results.cpu().numpy(), om[0])
    np.testing.assert_array_equal(cls_t.cpu().numpy() - 1, oc[0])
    np.testing.assert_allclose(reg_t.cpu().numpy(), ob[0], rtol=RTOL, atol=1e-7)
    sim = lab.target_assigner._similarity_calc.compare(BoxList(torch.from_numpy(gb[0]).to(dev())), BoxList(anc.boxes))
    np.testing.assert_array_equal(sim.cpu().numpy(), orc.iou_matrix(gb[0], anc.boxes.cpu().numpy()))
    with pytest.raises(ValueError):
        BoxList(torch.zeros(3, 5))
    with pytest.raises(ValueError):
        lab.target_assigner.assign(anc.boxes, BoxList(anc.boxes))


# ------------------------------------------------------------------------------------------ loss
def t(x, grad=False):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev()).requires_grad_(grad)


def assert_grad_close(got, ref, weight, ulps=4.0, what=''):
    """The gradient bar: 1e-5 relative (north_star) plus `ulps` units in the last place of 1.0 times the element
    weight.  d loss / d logit = weight * (sigmoid(x) - t) (times a modulator in the legacy focal form): where
    sigmoid(x) is within a few ulps of t the difference CANCELS, so two correctly rounded fp32 evaluations -- the
    reference's own autograd included -- differ by ~1 ulp(1) * weight there, far more than 1e-5 of a result that
    is itself ~0.  test_loss_gradients_vs_fp64 measures the actual errors against a float64 evaluation."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    bar = 1e-5 * np.abs(ref) + ulps * 2.0 ** -24 * weight
    err = np.abs(got - ref)
    bad = err > bar
    assert not bad.any(), f'{what}: {int(bad.sum())} gradient elements off; worst err/bar = {float((err / bar).max()):.2f}'


LOSS_TAGS = ['new_c90', 'new_c1', 'new_smooth', 'legacy', 'legacy_g0', 'new_256']


@pytest.mark.parametrize('tag', LOSS_TAGS)
def test_loss_golden(golden, tag):
    """loss_fn on reference-layout targets vs the reference's own values and autograd grads."""
    from ood_object_detection_b200.loss import loss_fn
    from test_oracle_golden import loss_case
    g = golden('loss')
    c = loss_case(g, tag)
    co, bo = [t(x, True) for x in c['co']], [t(x, True) for x in c['bo']]
    cls_t, box_t = [t(x) for x in c['cls_t']], [t(x) for x in c['box_t']]
    tot, cl, bl = loss_fn(co, bo, cls_t, box_t, t(c['npos']), c['C'], c['alpha'], c['gamma'], c['delta'], c['w'],
                          c['sm'], c['legacy'])
    np.testing.assert_allclose([tot.item(), cl.item(), bl.item()], g[f'{tag}_loss'], rtol=RTOL)
    tot.backward()
    _, _, _, ogc, ogb = orc.loss_fn(c['co'], c['bo'], c['cls_t'], c['box_t'], c['npos'], c['C'], c['alpha'], c['gamma'],
                                    c['delta'], c['w'], c['sm'], c['legacy'], want_grad=True)
    n = float(np.sum(c['npos'], dtype=np.float32)) + 1.0
    w_cls = max(c['alpha'], 1.0 - c['alpha']) / n * (1.0 + abs(c['gamma']) if c['legacy'] else 1.0)
    w_box = c['w'] / (4.0 * n)
    for l in range(5):
        assert_grad_close(bo[l].grad.cpu().numpy(), g[f'{tag}_gbox{l}'], w_box, what=f'box grad level {l} vs reference')
        assert_grad_close(co[l].grad.cpu().numpy(), ogc[l], w_cls, what=f'class grad level {l} vs oracle')
        if f'{tag}_gcls{l}' in g:
            assert_grad_close(co[l].grad.cpu().numpy(), g[f'{tag}_gcls{l}'], w_cls, what=f'class grad level {l} vs reference autograd')


@pytest.mark.parametrize('name,B,C,m,legacy,sm', [('d0', 4, 90, 10, False, 0.0), ('d0', 2, 20, 30, True, 0.0),
                                                  ('d0', 2, 7, 10, False, 0.1), ('d3', 2, 90, 20, False, 0.0),
                                                  ('d0', 3, 1, 10, False, 0.0)])
def test_loss_fused_vs_oracle(name, B, C, m, legacy, sm):
    """labeler + fused loss (targets never materialised) == oracle labeler + oracle loss, and
    == our own unfused path; D3's 7x7 level exercises the scalar (HW % 4 != 0) code path."""
    from ood_object_detection_b200.loss import loss_fn, loss_fn_fused
    size, scale = synth.MODEL_SHAPES[name]
    anc, lab = make_labeler(size, scale, C)
    gb, gc = synth.gt_boxes(300 + B + C, B, size, m, C)
    co_np, bo_np = synth.head_outputs(400 + B + C, B, size, C, tie_free=False)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0, label_smoothing=sm, legacy_focal=legacy)
    oc, ob, onp, _, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), list(gb), list(gc))
    fhw = synth.feat_hw(size)
    ref = orc.loss_fn(co_np, bo_np, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0,
                      sm, legacy, want_grad=True)
    co, bo = [t(x, True) for x in co_np], [t(x, True) for x in bo_np]
    lb = lab.assign(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()))
    tot, cl, bl = loss_fn_fused(co, bo, lb, **kw)
    np.testing.assert_allclose([tot.item(), cl.item(), bl.item()], ref[:3], rtol=RTOL)
    (tot * 2.0).backward()                      # upstream gradient != 1 exercises odk_scale_inplace
    n = float(np.sum(onp, dtype=np.float32)) + 1.0
    w_cls = 2.0 * 0.75 / n * (2.5 if legacy else 1.0)
    for l in range(5):
        assert_grad_close(co[l].grad.cpu().numpy(), 2.0 * ref[3][l], w_cls, what=f'class grad level {l}')
        assert_grad_close(bo[l].grad.cpu().numpy(), 2.0 * ref[4][l], 2.0 * 50.0 / (4.0 * n), what=f'box grad level {l}')
    # unfused path on materialised targets gives the same numbers
    cls_t, box_t = lb.targets()
    with torch.no_grad():
        tot2, cl2, bl2 = loss_fn([x.detach() for x in co], [x.detach() for x in bo], cls_t, box_t, lb.num_positives, **kw)
    np.testing.assert_allclose([tot2.item(), cl2.item(), bl2.item()], [tot.item(), cl.item(), bl.item()], rtol=1e-6)


def test_loss_gradients_vs_fp64():
    """How exact the fused kernel's gradients are: against the same loss evaluated in float64 (torch autograd on the
    device).  The kernel uses ex2.approx + a degree-7 log1p polynomial + a fast divide; its class gradients must
    stay inside the bar of assert_grad_close, and on the elements that do not cancel (|g| >= 1% of the weight)
    inside 1e-5 relative.  Prints the measured maxima (quoted in DESIGN.md)."""
    import torch.nn.functional as F
    from ood_object_detection_b200.loss import loss_fn_fused
    size, scale = synth.MODEL_SHAPES['d0']
    B, C, m, alpha, delta, w = 4, 90, 10, 0.25, 0.1, 50.0
    anc, lab = make_labeler(size, scale, C)
    gb, gc = synth.gt_boxes(77, B, size, m, C)
    co_np, bo_np = synth.head_outputs(78, B, size, C, tie_free=False)
    co, bo = [t(x, True) for x in co_np], [t(x, True) for x in bo_np]
    lb = lab.assign(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()))
    tot, _, _ = loss_fn_fused(co, bo, lb, num_classes=C, alpha=alpha, gamma=1.5, delta=delta, box_loss_weight=w)
    tot.backward()
    cls_t, box_t = lb.targets()
    n = lb.num_positives.double().sum() + 1.0
    co64 = [x.detach().double().requires_grad_(True) for x in co]
    bo64 = [x.detach().double().requires_grad_(True) for x in bo]
    total64 = 0.0
    for l in range(5):
        Bq, _, H, W = co64[l].shape
        x = co64[l].permute(0, 2, 3, 1).reshape(Bq, H, W, 9, C)
        ct = cls_t[l]
        oh = F.one_hot(ct.clamp(min=0), C).double() * (ct >= 0).unsqueeze(-1)
        bce = F.binary_cross_entropy_with_logits(x, oh, reduction='none')
        at = oh * alpha + (1.0 - oh) * (1.0 - alpha)
        cls_loss = (at * bce / n * (ct != -2).unsqueeze(-1)).sum()
        o, tg = bo64[l].permute(0, 2, 3, 1), box_t[l].double()
        ae = (o - tg).abs()
        q = ae.clamp(max=delta)
        box_loss = ((0.5 * q * q + delta * (ae - q)) * (tg != 0.0)).sum() / (n * 4.0)
        total64 = total64 + cls_loss + w * box_loss
    total64.backward()
    assert abs(tot.detach().item() - float(total64)) <= 1e-5 * abs(float(total64))
    w_cls = max(alpha, 1.0 - alpha) / float(n)
    worst_rel = worst_bar = 0.0
    for l in range(5):
        g, r = co[l].grad.double(), co64[l].grad
        err = (g - r).abs()
        big = r.abs() >= 1e-2 * w_cls
        worst_rel = max(worst_rel, float((err[big] / r.abs()[big]).max()))
        worst_bar = max(worst_bar, float((err / (1e-5 * r.abs() + 4 * 2.0 ** -24 * w_cls)).max()))
        assert_grad_close(bo[l].grad.cpu().numpy(), bo64[l].grad.cpu().numpy(), w / (4.0 * float(n)), what=f'box grad level {l} vs fp64')
    print(f'class gradients vs float64: max relative error on non-cancelling elements {worst_rel:.2e}; '
          f'max error / bar {worst_bar:.2f}')
    assert worst_rel <= 1e-5 and worst_bar <= 1.0


def test_loss_requires_cuda():
    from ood_object_detection_b200.loss import loss_fn
    x = [torch.zeros(1, 9, 4, 4)]
    with pytest.raises(RuntimeError):
        loss_fn(x, [torch.zeros(1, 36, 4, 4)], [torch.zeros(1, 4, 4, 9, dtype=torch.long)], [torch.zeros(1, 4, 4, 36)],
                torch.ones(1), 1, 0.25, 1.5, 0.1, 50.0)


@pytest.mark.parametrize('legacy,sm', [(False, 0.0), (False, 0.1), (True, 0.0)])
def test_loss_channels_last_inputs_in_place(legacy, sm):
    """channels_last head outputs ([B, H, W, C] in memory) go through the fused labeler + loss
    without a layout copy (layout-agnostic stream + per-anchor patch kernels): same losses and gradients as the
    NCHW path and the oracle, gradients returned in the inputs' own memory format; mixed layouts per level too
    (D3: odd 7x7 level)."""
    from ood_object_detection_b200.loss import loss_fn_fused
    size, scale = synth.MODEL_SHAPES['d3']
    B, C, m = 2, 90, 12
    anc, lab = make_labeler(size, scale, C)
    gb, gc = synth.gt_boxes(501, B, size, m, C)
    co_np, bo_np = synth.head_outputs(502, B, size, C, tie_free=False)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0, label_smoothing=sm, legacy_focal=legacy)
    oc, ob, onp, _, _ = orc.batch_label_anchors(anc.boxes.cpu().numpy(), list(gb), list(gc))
    fhw = synth.feat_hw(size)
    ref = orc.loss_fn(co_np, bo_np, orc.split_levels(oc, fhw), orc.split_levels(ob, fhw), onp, C, 0.25, 1.5, 0.1, 50.0, sm, legacy, want_grad=True)
    n = float(np.sum(onp, dtype=np.float32)) + 1.0
    w_cls = 0.75 / n * (2.5 if legacy else 1.0)
    for pattern in ('all', 'mixed'):
        def fmt(x, l, which):
            cl = pattern == 'all' or (l + which) % 2 == 0
            tt = t(x)
            return (tt.contiguous(memory_format=torch.channels_last) if cl else tt).requires_grad_(True)
        co = [fmt(x, l, 0) for l, x in enumerate(co_np)]
        bo = [fmt(x, l, 1) for l, x in enumerate(bo_np)]
        ptrs = [x.data_ptr() for x in co]
        lb = lab.assign(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()), transient=True)
        tot, cl_, bl_ = loss_fn_fused(co, bo, lb, **kw)
        np.testing.assert_allclose([tot.item(), cl_.item(), bl_.item()], ref[:3], rtol=RTOL)
        tot.backward()
        assert [x.data_ptr() for x in co] == ptrs
        for l in range(5):
            assert co[l].grad.stride() == co[l].stride() and bo[l].grad.stride() == bo[l].stride()
            assert_grad_close(co[l].grad.cpu().numpy(), ref[3][l], w_cls, what=f'{pattern}: class grad level {l}')
            assert_grad_close(bo[l].grad.cpu().numpy(), ref[4][l], 50.0 / (4.0 * n), what=f'{pattern}: box grad level {l}')
    # the labeler's workspace came back zeroed from the transient batch: the next assignment is right without a memset
    lb2 = lab.assign(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()))
    np.testing.assert_array_equal(lb2.num_positives.cpu().numpy(), onp)


@pytest.mark.parametrize('name,B,m', [('d0', 4, 10), ('d3', 2, 60)])
def test_loss_stream_patch_modes_bitwise(name, B, m):
    """The fused loss = a layout-agnostic stream over the logits + a patch of the matched anchors.  The patch either
    walks the labeler's list of matched anchors (transient label batch: atomic-append order, and it zeroes the keys)
    or scans the key / match row; its sums are fixed-point integers, so both modes, and repeated runs on lists in
    different orders, give the SAME BITS -- losses and gradients.  The workspace is left clean for the next step."""
    from ood_object_detection_b200.loss import loss_fn_fused
    size, scale = synth.MODEL_SHAPES[name]
    C = 90
    anc, lab = make_labeler(size, scale, C)
    gb, gc = synth.gt_boxes(910 + B, B, size, m, C)
    co_np, bo_np = synth.head_outputs(911 + B, B, size, C, tie_free=False)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)
    gbt, gct = torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev())

    def run(transient):
        co, bo = [t(x, True) for x in co_np], [t(x, True) for x in bo_np]
        lb = lab.assign(gbt, gct, transient=transient)
        tot, cl, bl = loss_fn_fused(co, bo, lb, **kw)
        tot.backward()
        return (np.array([tot.item(), cl.item(), bl.item()], np.float32), [x.grad.cpu().numpy() for x in co],
                [x.grad.cpu().numpy() for x in bo])
    ref = run(False)
    for trial in range(3 if USE_GRID[0] else 1):   # (the list exists on the grid labeler only)
        got = run(USE_GRID[0])
        np.testing.assert_array_equal(got[0].view(np.uint32), ref[0].view(np.uint32))
        for l in range(5):
            np.testing.assert_array_equal(got[1][l], ref[1][l])
            np.testing.assert_array_equal(got[2][l], ref[2][l])


def test_loss_stream_agrees_with_plane_kernels(monkeypatch):
    """ODK_LOSS_KERNEL=ring selects the plane-walking kernels (the round-1 path, still used for targets given as
    tensors), ODK_LOSS_TMA=0 the LDG/STG tile stream for the gradient pass instead of the TMA bulk tiles: same losses
    to 1e-6 (different summation order) and the same gradients bit for bit."""
    from ood_object_detection_b200.loss import loss_fn_fused
    size, scale = synth.MODEL_SHAPES['d3']
    B, C, m = 2, 90, 25
    anc, lab = make_labeler(size, scale, C)
    gb, gc = synth.gt_boxes(920, B, size, m, C)
    co_np, bo_np = synth.head_outputs(921, B, size, C, tie_free=False)
    kw = dict(num_classes=C, alpha=0.25, gamma=1.5, delta=0.1, box_loss_weight=50.0)
    out = {}
    for which in ('stream', 'stream_ldg', 'ring'):   # TMA bulk tiles (default) | LDG/STG tiles | plane-walking kernels
        monkeypatch.delenv('ODK_LOSS_KERNEL', raising=False)
        monkeypatch.delenv('ODK_LOSS_TMA', raising=False)
        if which == 'ring':
            monkeypatch.setenv('ODK_LOSS_KERNEL', 'ring')
        elif which == 'stream_ldg':
            monkeypatch.setenv('ODK_LOSS_TMA', '0')
        co, bo = [t(x, True) for x in co_np], [t(x, True) for x in bo_np]
        lb = lab.assign(torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev()))
        tot, cl, bl = loss_fn_fused(co, bo, lb, **kw)
        tot.backward()
        out[which] = ([tot.item(), cl.item(), bl.item()], [x.grad.cpu().numpy() for x in co], [x.grad.cpu().numpy() for x in bo])
    for other in ('stream_ldg', 'ring'):
        np.testing.assert_allclose(out['stream'][0], out[other][0], rtol=1e-6)
        for l in range(5):
            np.testing.assert_array_equal(out['stream'][1][l], out[other][1][l])
            np.testing.assert_array_equal(out['stream'][2][l], out[other][2][l])


def test_anchor_table_from_generator():
    """odk_anchor_table: the labeler kernel's recomputed anchors (float64 generator -> fp32) against the table, every
    anchor of every model shape, bit for bit -- the precondition for use_anchor_generator to leave assignments
    untouched (they are additionally compared with the oracle's in every labeler test, which run with it on)."""
    from ood_object_detection_b200 import _lib
    from ood_object_detection_b200.anchors import Anchors
    lib = _lib.lib()
    for name, (size, scale) in synth.MODEL_SHAPES.items():
        anc = Anchors(3, 7, 3, synth.ASPECTS, scale, (size, size)).to(dev())
        out = torch.empty_like(anc.boxes)
        hw = anc.level_hw()
        with torch.cuda.device(dev()):
            _lib.check(lib.odk_anchor_table(_lib.ptr(anc.plane_desc), _lib.ptr(anc.plane_gen), anc.plane_desc.shape[0], _lib.int_array(hw),
                                            len(hw), anc.get_anchors_per_location(), _lib.ptr(out), _lib.stream_ptr(dev())))
        assert torch.equal(out, anc.boxes), name


def test_labeler_gather_and_generator_paths_agree():
    """The same assignment with the anchors gathered from the table (use_anchor_generator = False) and recomputed."""
    size, scale = synth.MODEL_SHAPES['d3']
    anc, lab = make_labeler(size, scale, 90)
    gb, gc = synth.gt_boxes(931, 3, size, 40, 90)
    gbt, gct = torch.from_numpy(gb).to(dev()), torch.from_numpy(gc).to(dev())
    res = []
    for use in (True, False):
        lab.use_anchor_generator = use
        lb = lab.assign(gbt, gct)
        res.append((lb.match.clone(), lb.num_positives.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])

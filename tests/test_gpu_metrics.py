"""mIoU / PQ kernels vs the reference's known-answer tests, the golden vectors and the
oracle.  float64 PQ states are compared BIT-exactly."""
import numpy as np
import pytest
import torch

import oracle
from conftest import jload, load_golden

pytestmark = pytest.mark.gpu


def _pq(dev, **kw):
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    return PanopticQuality(device=dev, **kw)


def _states(m):
    return [getattr(m, n).cpu().numpy() for n in
            ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')]


# ---- the reference's known-answer cases (tests/test_metrics.py:76-446) ----------------------
INST_A = [[1, 1, 1, 1, 1, 1], [1, 2, 2, 2, 2, 1], [1, 2, 2, 2, 2, 1],
          [1, 2, 2, 2, 2, 1], [1, 2, 2, 1, 1, 1], [1, 2, 1, 1, 1, 1]]
GT_B = [[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 2, 2, 2, 1],
        [1, 2, 2, 2, 2, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]]
GOOD_B = [[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 2, 2, 2, 2, 1],
          [1, 2, 2, 2, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]]
BAD_B = [[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 2, 2, 1],
         [1, 1, 1, 2, 2, 1], [1, 1, 1, 2, 2, 1], [1, 1, 1, 1, 1, 1]]
CAT_C = [[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 2, 2, 1, 2, 2],
         [1, 2, 2, 1, 2, 2], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]]


def _t(a, dev):
    return torch.tensor(a, dtype=torch.int64, device=dev)[None]


def test_pq_perfect_match(cuda_device):
    m = _pq(cuda_device, num_categories=1, ignored_label=2, max_instances_per_category=16,
            offset=16, is_thing=torch.tensor([True]))
    x = _t(INST_A, cuda_device)
    m.update(x, x)
    iou, tp, fn, fp = _states(m)
    assert (iou, tp, fn, fp) == ([2.0], [2], [0], [0])
    r = m.compute()
    assert r['pq_per_class'].tolist() == [1.0] and r['rq_per_class'].tolist() == [1.0]
    assert r['all_pq'] == 1.0 and r['all_rq'] == 1.0 and r['all_sq'] == 1.0
    assert r['all_num_categories'] == 1


def test_pq_totally_wrong(cuda_device):
    cat = torch.tensor([[0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 1, 0], [0, 1, 1, 1, 1, 0],
                        [0, 1, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]],
                       dtype=torch.int64, device=cuda_device)[None]
    m = _pq(cuda_device, num_categories=2, ignored_label=2, max_instances_per_category=1,
            offset=16, is_thing=torch.tensor([True, True]))
    m.update(1 - cat, cat)
    iou, tp, fn, fp = _states(m)
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == ([0, 0], [0, 0], [1, 1], [1, 1])
    r = m.compute()
    assert r['all_pq'] == 0.0 and r['all_num_categories'] == 2


def test_pq_matches_by_iou(cuda_device):
    m = _pq(cuda_device, num_categories=1, ignored_label=2, max_instances_per_category=16,
            offset=16, is_thing=torch.tensor([True]))
    m.update(_t(GOOD_B, cuda_device), _t(GT_B, cuda_device))
    iou, tp, fn, fp = _states(m)
    assert iou[0] == 28 / 30 + 6 / 8 and (tp, fn, fp) == ([2], [0], [0])
    r = m.compute()
    assert r['pq_per_class'].tolist() == [(28 / 30 + 6 / 8) / 2]
    assert r['all_sq'] == (28 / 30 + 6 / 8) / 2 and r['all_rq'] == 1.0
    m.reset()
    m.update(_t(BAD_B, cuda_device), _t(GT_B, cuda_device))
    iou, tp, fn, fp = _states(m)
    assert iou[0] == 27 / 32 and (tp, fn, fp) == ([1], [1], [1])
    r = m.compute()
    assert r['all_pq'] == 27 / 32 / 2 and r['all_rq'] == 0.5 and r['all_sq'] == 27 / 32


def test_pq_wrong_instances_and_arbitrary_order(cuda_device):
    cat = torch.tensor(CAT_C, dtype=torch.int64, device=cuda_device)[None]
    pinst = torch.zeros_like(cat)
    pinst[:, 2:4, 4:6] = 1
    kw = dict(num_categories=3, ignored_label=0, max_instances_per_category=10, offset=100,
              is_thing=[True, True, True])
    m = _pq(cuda_device, **kw)
    m.update(cat * 10 + pinst, cat * 10)
    iou, tp, fn, fp = _states(m)
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == \
        ([0, 1, 0], [0, 1, 0], [0, 0, 1], [0, 0, 2])
    r = m.compute()
    assert r['all_pq'] == 0.5 and r['all_num_categories'] == 2
    ginst = torch.zeros_like(cat)
    ginst[:, 2:4, 1:3] = 1
    m = _pq(cuda_device, **kw)
    m.update(cat * 10 + pinst, cat * 10 + ginst)
    iou, tp, fn, fp = _states(m)
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == \
        ([0, 1, 2], [0, 1, 2], [0, 0, 0], [0, 0, 0])
    assert m.compute()['all_pq'] == 1.0


def test_pq_multiple_batches(cuda_device):
    m = _pq(cuda_device, num_categories=1, ignored_label=2, max_instances_per_category=16,
            offset=16, is_thing=[True])
    gt = torch.cat((_t(GT_B, cuda_device),) * 2)
    m.update(gt, torch.cat((_t(GOOD_B, cuda_device),) * 2))
    m.update(gt, torch.cat((_t(BAD_B, cuda_device),) * 2))
    r = m.compute()
    assert r['pq_per_class'].tolist() == [((28 / 30 + 6 / 8) + (27 / 32)) / 2 / 2]
    assert r['rq_per_class'].tolist() == [3 / 4]
    assert r['sq_per_class'].tolist() == [((28 / 30 + 6 / 8) + (27 / 32)) / 3]
    assert r['all_rq'] == 0.75 and r['all_num_categories'] == 1


# ---- golden vectors / oracle -----------------------------------------------------------------
def test_pq_golden_frames(cuda_device):
    from nicr_mt_scene_analysis_b200.metric import compare_and_accumulate
    z = load_golden('pq')
    NC, L, OFF = int(z['num_categories']), int(z['L']), int(z['offset'])
    pred = torch.from_numpy(z['pred']).to(cuda_device)
    tgt = torch.from_numpy(z['target']).to(cuda_device)
    matches = jload(z['matches'])
    for b in range(pred.shape[0]):
        iou, tp, fn, fp, m = compare_and_accumulate(pred[b], tgt[b], NC, 0, L, OFF, 0)
        assert iou.dtype == torch.float64
        assert np.array_equal(iou.numpy(), z['iou'][b])          # bit-exact float64
        assert np.array_equal(tp.numpy(), z['tp'][b])
        assert np.array_equal(fn.numpy(), z['fn'][b])
        assert np.array_equal(fp.numpy(), z['fp'][b])
        assert sorted([list(x) for x in m]) == matches[b]
    pq = _pq(cuda_device, num_categories=NC, ignored_label=0, max_instances_per_category=L,
             offset=OFF, is_thing=z['is_thing'].tolist())
    pq.update(pred[:4], tgt[:4])
    pq.update(pred[4:], tgt[4:])
    assert np.array_equal(np.stack(_states(pq)), z['state'])   # frame-order accumulation
    res = pq.compute()
    ref = oracle.pq_results(*z['state'], z['is_thing'], 0)
    for k, v in ref.items():
        np.testing.assert_allclose(np.asarray(res[k], dtype=np.float64), v, rtol=1e-12)


def test_pq_zero_division_raises(cuda_device):
    m = _pq(cuda_device, num_categories=2, ignored_label=0, max_instances_per_category=10,
            offset=100, is_thing=[True, True])
    m.update(torch.full((1, 4, 4), 3, dtype=torch.int64, device=cuda_device),
             torch.zeros((1, 4, 4), dtype=torch.int64, device=cuda_device))
    with pytest.raises(ZeroDivisionError):
        m.compute()


def test_pq_wide_target_ids_standard_geometry(cuda_device):
    """Standard id geometry (offset 256^3, 65536 instances / category) with ground-truth ids
    beyond 32 bits: those lanes leave the 32-bit fast path of the pixel pass; the result still
    equals the oracle's int64 arithmetic.  The wide ids belong to the ignored category."""
    from nicr_mt_scene_analysis_b200.metric import compare_and_accumulate
    L, OFF, NC, IGN = 1 << 16, 256 ** 3, 5, 1 << 16
    g = torch.Generator().manual_seed(7)
    H, W = 64, 96
    cat = torch.randint(1, NC, (H // 8, W // 8), generator=g).repeat_interleave(8, 0).repeat_interleave(8, 1)
    inst = torch.randint(0, 3, (H // 4, W // 4), generator=g).repeat_interleave(4, 0).repeat_interleave(4, 1)
    tgt = cat * L + inst
    pred = torch.roll(tgt, 2, 1).clone()
    tgt[10:30, 17:43] = IGN * L + 3          # 2^32 + 3: category == ignored_label
    tgt[40:44, 1:3] = IGN * L                # a second wide segment, not aligned to 4 pixels
    ref = oracle.pq_compare_and_accumulate(pred.numpy(), tgt.numpy(), NC, IGN, L, OFF, 0)
    iou, tp, fn, fp, m = compare_and_accumulate(pred.to(cuda_device), tgt.to(cuda_device), NC, IGN,
                                                L, OFF, 0)
    for got, want in zip((iou, tp, fn, fp), ref[:4]):
        assert np.array_equal(got.numpy(), want)
    assert sorted([list(x) for x in m]) == sorted([list(x) for x in ref[4]])


def _many_pairs_frame(seed, n_gt, n_pred, side=60):
    """side*side pixels, (nearly) every one a different (gt segment, pred segment) pair."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(side * side, generator=g).reshape(side, side)
    L = 1 << 16
    tgt = 1 * L + (idx % n_gt) + 1
    pred = 1 * L + (idx % n_pred) + 1
    return pred, tgt


@pytest.mark.parametrize('frames', [2, 600])
def test_pq_thousands_of_pairs_per_frame(frames, cuda_device):
    """~3400 distinct pairs in a 3600-pixel frame (997 gt x 241 pred segments): far beyond the
    2048-slot table of a CTA of the pixel pass.  With 600 frames in the batch one CTA owns a whole
    frame, so its table overflows and single entries go straight to the frame's list."""
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    L, OFF, NC = 1 << 16, 256 ** 3, 3
    distinct = [_many_pairs_frame(s, 997, 241) for s in range(3)]
    pred = torch.stack([distinct[b % 3][0] for b in range(frames)])
    tgt = torch.stack([distinct[b % 3][1] for b in range(frames)])
    pq = PanopticQuality(NC, 0, L, OFF, [False, True, True], device=cuda_device)
    pq.update(pred.to(cuda_device), tgt.to(cuda_device))
    pq.check_status()
    per_frame = [oracle.pq_compare_and_accumulate(p.numpy(), t.numpy(), NC, 0, L, OFF, 0)[:4]
                 for p, t in distinct]
    state = np.zeros((4, NC))
    for b in range(frames):
        for s, v in zip(state, per_frame[b % 3]):
            s += v
    assert np.array_equal(np.stack(_states(pq)), state)
    assert state[2].sum() > 0 and state[3].sum() > 0        # the case is not trivial


def test_pq_large_frame_path(cuda_device):
    """A frame beyond the shared-memory matcher (~6300 distinct pairs, 1499 gt / 1013 pred
    segments) between two ordinary ones: it contributes nothing in the batched pass and is
    evaluated again with the global-memory tables and put back in its place: the states equal
    the oracle's bit for bit, the float64 IoU sums included (frame order of pq.py:298-303)."""
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality, compare_and_accumulate)
    L, OFF, NC = 1 << 16, 256 ** 3, 3
    small = _many_pairs_frame(1, 31, 17, side=80)
    big = _many_pairs_frame(5, 1499, 1013, side=80)
    pred = torch.stack([small[0], big[0], small[0]])
    tgt = torch.stack([small[1], big[1], small[1]])
    pq = PanopticQuality(NC, 0, L, OFF, [False, True, True], device=cuda_device)
    miou = MeanIntersectionOverUnion(NC, True, device=cuda_device)
    sem_t = (tgt // L).to(torch.uint8)
    PanopticEvaluation(pq, miou).update(pred.to(cuda_device), tgt.to(cuda_device), sem_t.to(cuda_device))
    pq.update(pred.to(cuda_device), tgt.to(cuda_device))      # both updates are followed up below
    pq.check_status()
    state = np.zeros((4, NC))
    frames = [oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), NC, 0, L, OFF, 0)
              for b in range(3)]
    for _ in range(2):
        for out in frames:
            for s, v in zip(state, out[:4]):
                s += v
    got = np.stack(_states(pq))
    assert np.array_equal(got, state)       # tp / fn / fp AND the IoU sums, in frame order
    assert np.array_equal(miou.confmat.cpu().numpy(), oracle.confmat((pred // L).numpy(), sem_t.numpy(), NC))
    # single-frame API: same path, same float64 order as the reference -> bit-exact, with matches
    iou, tp, fn, fp, m = compare_and_accumulate(big[0].to(cuda_device), big[1].to(cuda_device),
                                                NC, 0, L, OFF, 0)
    for a, b in zip((iou, tp, fn, fp), frames[1][:4]):
        assert np.array_equal(a.numpy(), b)
    assert m == frames[1][4]


def test_pq_large_frames_keep_the_frame_order_over_several_updates(cuda_device):
    """Seven eager updates of three frames, over-capacity frames in the second, fourth (two of
    them) and last one: every follow-up rebuilds the states from the journal of the oldest
    affected update, so the float64 IoU sums equal the frame-by-frame order of the reference
    (pq.py:298-303) bit for bit -- also when another large frame turns up among the updates that
    are re-accumulated, and with ordinary updates in between."""
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    L, OFF, NC = 1 << 16, 256 ** 3, 3
    def frame(seed, shift, noisy_rows):
        # 16x16 blocks of alternating classes, the prediction shifted by `shift` px (matches with
        # IoUs (16 - shift) / (16 + shift) and clipped ones at the border); the upper
        # `noisy_rows` rows carry a different (gt, pred) pair in every pixel
        side = 96
        yy, xx = torch.meshgrid(torch.arange(side), torch.arange(side), indexing='ij')
        blk = (yy // 16) * 6 + xx // 16
        tgt = (1 + blk % 2) * L + blk + 1
        pred = torch.roll(tgt, shifts=(shift, seed % 3), dims=(1, 0))
        if noisy_rows:
            g = torch.Generator().manual_seed(seed)
            n = noisy_rows * side
            idx = torch.randperm(n, generator=g).reshape(noisy_rows, side)
            tgt[:noisy_rows] = 1 * L + 100 + idx % 1499
            pred[:noisy_rows] = 1 * L + 100 + idx % 1013
        return pred, tgt
    smalls = [frame(s, s, 0) for s in range(1, 6)]
    bigs = [frame(7, 2, 64), frame(9, 4, 80)]
    plan = [(0, 1, 2), (3, 'b0', 4), (1, 1, 0), ('b1', 2, 'b0'), (4, 3, 2), (0, 0, 1), (2, 'b1', 3)]
    pick = lambda k: bigs[int(k[1])] if isinstance(k, str) else smalls[k]
    pq = PanopticQuality(NC, 0, L, OFF, [False, True, True], device=cuda_device)
    state = np.zeros((4, NC))
    cache = {}
    for upd in plan:
        pred = torch.stack([pick(k)[0] for k in upd])
        tgt = torch.stack([pick(k)[1] for k in upd])
        pq.update(pred.to(cuda_device), tgt.to(cuda_device))
        for k in upd:
            if k not in cache:
                cache[k] = oracle.pq_compare_and_accumulate(pick(k)[0].numpy(), pick(k)[1].numpy(),
                                                            NC, 0, L, OFF, 0)[:4]
            for s, v in zip(state, cache[k]):
                s += v
    pq.check_status()
    assert state[0].sum() != np.floor(state[0].sum())      # (the sums are not trivially exact)
    assert np.array_equal(np.stack(_states(pq)), state)
    # and once more on top of the non-zero states
    for upd in plan[:4]:
        pred = torch.stack([pick(k)[0] for k in upd])
        tgt = torch.stack([pick(k)[1] for k in upd])
        pq.update(pred.to(cuda_device), tgt.to(cuda_device))
        for k in upd:
            for s, v in zip(state, cache[k]):
                s += v
    pq.check_status()
    assert np.array_equal(np.stack(_states(pq)), state)


def test_pq_capacity_error_inside_a_cuda_graph(cuda_device):
    """Updates replayed from a CUDA graph cannot be followed up: an overflowing frame stays an
    error there (never a silently wrong result)."""
    from nicr_mt_scene_analysis_b200._lib import ERR_CAPACITY, NpbError
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    L, OFF = 1 << 16, 256 ** 3
    pred, tgt = _many_pairs_frame(5, 1499, 1013, side=80)
    pred, tgt = pred[None].to(cuda_device), tgt[None].to(cuda_device)
    pq = PanopticQuality(3, 0, L, OFF, [False, True, True], device=cuda_device)
    step = CapturedStep(lambda: pq.update(pred, tgt), warmup=1, device=cuda_device)
    pq.reset()
    step.replay()
    with pytest.raises(NpbError) as err:
        pq.compute()
    assert err.value.code == ERR_CAPACITY


@pytest.mark.parametrize('bad', ['pred_negative', 'pred_ge_offset', 'target_negative'])
def test_pq_id_range_errors_standard_geometry(bad, cuda_device):
    from nicr_mt_scene_analysis_b200._lib import NpbError
    m = _pq(cuda_device, num_categories=4, ignored_label=0, max_instances_per_category=1 << 16,
            offset=256 ** 3, is_thing=[False, True, True, False])
    pred = torch.full((1, 8, 16), (1 << 16) + 1, dtype=torch.int64)
    tgt = pred.clone()
    if bad == 'pred_negative':
        pred[0, 3, 5] = -1
    elif bad == 'pred_ge_offset':
        pred[0, 3, 5] = 256 ** 3
    else:
        tgt[0, 7, 15] = -5
    m.update(pred.to(cuda_device), tgt.to(cuda_device))
    with pytest.raises(NpbError):
        m.compute()


@pytest.mark.parametrize('n', [6, 41, 200])
def test_miou_golden(n, cuda_device):
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion
    z = load_golden('miou')
    for flag in (False, True):
        m = MeanIntersectionOverUnion(n_classes=n, ignore_first_class=flag, device=cuda_device)
        for p, t in zip(z[f'pred_{n}'], z[f'target_{n}']):
            tt = torch.from_numpy(t)
            m.update(torch.from_numpy(p).to(cuda_device),
                     (tt.to(torch.uint8) if n <= 255 and not flag else tt).to(cuda_device))
        assert m.confmat.dtype == torch.int64
        assert np.array_equal(m.confmat.cpu().numpy(), z[f'confmat_{n}'])
        miou, ious = m.compute(return_ious=True)
        assert miou.numpy() == z[f'miou{int(flag)}_{n}']
        assert np.array_equal(ious.numpy(), z[f'ious{int(flag)}_{n}'], equal_nan=True)
        m.reset()
        assert int(m.confmat.sum()) == 0


def test_miou_out_of_range_raises(cuda_device):
    from nicr_mt_scene_analysis_b200 import _lib
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion
    m = MeanIntersectionOverUnion(n_classes=4, device=cuda_device)
    m.update(torch.tensor([0, 1, 7], device=cuda_device), torch.tensor([0, 1, 2], device=cuda_device))
    with pytest.raises(_lib.NpbError):
        m.compute()


_TORCH_INT = {'u8': torch.uint8, 'i16': torch.int16, 'i32': torch.int32, 'i64': torch.int64}


@pytest.mark.parametrize('pd', ['u8', 'i16', 'i32', 'i64'])
@pytest.mark.parametrize('td', ['u8', 'i16', 'i32', 'i64'])
def test_miou_dtypes_alignment_and_tails(pd, td, cuda_device):
    """Every dtype pair through the streaming kernel (aligned maps, whole and partial groups of
    16 elements) and the generic kernel (maps that start at an odd element; n > 96), with and
    without the void skip, against the oracle's confusion matrix."""
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion
    g = torch.Generator().manual_seed(len(pd) * 7 + len(td))
    for n, N in ((5, 1), (19, 15), (41, 16), (41, 4096 + 17), (96, 3 * 4096 * 16 + 5), (97, 70001)):
        low = torch.randint(0, n + 1, (N // 50 + 2,), generator=g)
        target = low.repeat_interleave(50)[:N + 1].contiguous()           # runs of equal labels
        clean = (target - 1).clamp(min=0)
        preds = torch.where(torch.rand(N + 1, generator=g) < 0.2,
                            torch.randint(0, n, (N + 1,), generator=g), clean)
        p_dev = preds.to(_TORCH_INT[pd]).to(cuda_device)
        t_dev = target.to(_TORCH_INT[td]).to(cuda_device)
        for first in (0, 1):                     # first = 1: the maps start at an odd element
            p, t = preds[first:first + N], target[first:first + N]
            keep = t != 0
            m = MeanIntersectionOverUnion(n_classes=n, device=cuda_device)
            m.update_nonvoid(p_dev[first:first + N], t_dev[first:first + N])
            m.check_status()
            assert np.array_equal(m.confmat.cpu().numpy(),
                                  oracle.confmat(p[keep].numpy(), (t[keep] - 1).numpy(), n)), (n, N, first)
            m = MeanIntersectionOverUnion(n_classes=n + 1, device=cuda_device)
            m.update(p_dev[first:first + N], t_dev[first:first + N])
            m.update(p_dev[first:first + N], t_dev[first:first + N])      # accumulates
            m.check_status()
            assert np.array_equal(m.confmat.cpu().numpy(),
                                  2 * oracle.confmat(p.numpy(), t.numpy(), n + 1)), (n, N, first)


def test_miou_nonvoid_full_size_and_range(cuda_device):
    """480x640 x 8 frames: sum of the matrix == number of non-void pixels, rows / columns ==
    bincounts; predictions at void pixels are not range checked, the others are."""
    from nicr_mt_scene_analysis_b200 import _lib
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion
    C = 40
    g = torch.Generator().manual_seed(3)
    low = torch.randint(0, C + 1, (8, 15, 20), generator=g)
    target = low.repeat_interleave(32, 1).repeat_interleave(32, 2).to(torch.uint8).to(cuda_device)
    preds = torch.randint(0, C, (8, 480, 640), generator=g).to(cuda_device)
    preds[target == 0] = 1000                                    # garbage below void
    m = MeanIntersectionOverUnion(n_classes=C, device=cuda_device)
    m.update_nonvoid(preds, target)
    m.check_status()
    cm = m.confmat.cpu()
    keep = (target != 0).cpu()
    assert int(cm.sum()) == int(keep.sum())
    assert torch.equal(cm.sum(1), torch.bincount(target.cpu()[keep].long() - 1, minlength=C))
    assert torch.equal(cm.sum(0), torch.bincount(preds.cpu()[keep], minlength=C))
    preds[0, 0, :16] = C                                          # out of range at non-void pixels
    target[0, 0, :16] = 3
    m.update_nonvoid(preds, target)
    with pytest.raises(_lib.NpbError):
        m.compute()
    m = MeanIntersectionOverUnion(n_classes=C, device=cuda_device)
    target[0, 0, :16] = C + 1                                     # target - 1 out of range
    preds[0, 0, :16] = 0
    m.update_nonvoid(preds, target)
    with pytest.raises(_lib.NpbError):
        m.compute()


@pytest.mark.parametrize('shape', [(3, 120, 160), (2, 97, 131), (2, 530, 730)])
def test_fused_evaluation_against_oracle(shape, cuda_device):
    """PanopticEvaluation (one pixel pass) == separate PQ + mIoU updates == oracle."""
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    B, H, W = shape
    C, L, OFF = 12, 1 << 16, 256 ** 3
    g = torch.Generator().manual_seed(B * H)
    def blocky(n, blk):
        low = torch.randint(0, n, (B, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
        return low.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()
    is_thing = torch.tensor([False] + [bool(c % 2) for c in range(C)])
    cat_t, inst_t = blocky(C + 1, 24), blocky(4, 10)
    cat_p = torch.where(torch.rand(B, H, W, generator=g) < 0.1, blocky(C + 1, 24), cat_t)
    tgt = cat_t * L + torch.where(is_thing[cat_t], inst_t + 1, 0)
    pred = cat_p * L + torch.where(is_thing[cat_p], torch.roll(inst_t, 3, -1) + 1, 0)
    sem_t = cat_t.to(torch.uint8)
    kw = dict(num_categories=C + 1, ignored_label=0, max_instances_per_category=L, offset=OFF,
              is_thing=is_thing.tolist(), device=cuda_device)
    pq_f, pq_s = PanopticQuality(**kw), PanopticQuality(**kw)
    mi_f = MeanIntersectionOverUnion(C + 1, True, device=cuda_device)
    mi_s = MeanIntersectionOverUnion(C + 1, True, device=cuda_device)
    d = lambda t: t.to(cuda_device)
    PanopticEvaluation(pq_f, mi_f).update(d(pred), d(tgt), d(sem_t))
    pq_s.update(d(pred), d(tgt))
    mi_s.update(d(pred) // L, d(sem_t))
    pq_f.check_status()
    state = np.zeros((4, C + 1))
    for b in range(B):
        out = oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), C + 1, 0, L, OFF, 0)
        for s, v in zip(state, out[:4]):
            s += v
    assert np.array_equal(np.stack(_states(pq_f)), state)
    assert np.array_equal(np.stack(_states(pq_s)), state)
    cm = oracle.confmat((pred // L).numpy(), sem_t.numpy(), C + 1)
    assert np.array_equal(mi_f.confmat.cpu().numpy(), cm)
    assert np.array_equal(mi_s.confmat.cpu().numpy(), cm)
    # checksum property at any size: every pixel lands in exactly one confusion-matrix cell
    assert int(mi_f.confmat.sum()) == B * H * W


def test_pq_with_orientation_mae(cuda_device):
    from nicr_mt_scene_analysis_b200.metric import PanopticQualityWithOrientationMAE
    L = 1 << 16
    tgt = torch.zeros((1, 8, 8), dtype=torch.int64)
    tgt[:, :4] = 1 * L + 1
    tgt[:, 4:] = 1 * L + 2
    m = PanopticQualityWithOrientationMAE(num_categories=2, ignored_label=0,
                                          max_instances_per_category=L, offset=256 ** 3,
                                          is_thing=[False, True], device=cuda_device)
    m.update(tgt.to(cuda_device), [{7: 0.5, 9: 3.0}], [{1 * L + 1: 7, 1 * L + 2: 9}],
             tgt.to(cuda_device), [{3: 0.25, 4: -3.0}], [{1 * L + 1: 3, 1 * L + 2: 4}])
    r = m.compute(suffix='_deeplab')
    assert m.tp_per_class.tolist() == [0, 2]
    assert int(m.n_elements) == 2
    import math
    want = (0.25 + (2 * math.pi - 6.0)) / 2
    assert float(r['mae_deeplab_rad']) == pytest.approx(want, rel=1e-6)
    assert 'all_deeplab_pq' in r and 'mae_deeplab_deg' in r


@pytest.mark.parametrize('frames', [1, 8])
def test_pq_noise_ids_few_frames(frames, cuda_device):
    """Few frames per launch = many CTAs of the pixel pass per frame, and random ids = nearly
    every pixel of a CTA is its own pair (~5300 distinct pairs per frame: an untrained network's
    output looks like this).  The frames overflow the batched pass and take the large-frame path."""
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticEvaluation, PanopticQuality
    L, OFF, NC = 1 << 16, 256 ** 3, 9
    g = torch.Generator().manual_seed(100 + frames)
    H, W = 240, 320
    cat_t = torch.randint(0, NC, (frames, H, W), generator=g)
    cat_p = torch.randint(0, NC, (frames, H, W), generator=g)
    is_thing = [False] + [bool(c % 2) for c in range(1, NC)]
    it = torch.tensor(is_thing)
    tgt = cat_t * L + torch.where(it[cat_t], torch.randint(1, 8, (frames, H, W), generator=g), 0)
    pred = cat_p * L + torch.where(it[cat_p], torch.randint(1, 40, (frames, H, W), generator=g), 0)
    pq = PanopticQuality(NC, 0, L, OFF, is_thing, device=cuda_device)
    miou = MeanIntersectionOverUnion(NC, True, device=cuda_device)
    PanopticEvaluation(pq, miou).update(pred.to(cuda_device), tgt.to(cuda_device),
                                        cat_t.to(torch.uint8).to(cuda_device))
    pq.check_status()
    state = np.zeros((4, NC))
    for b in range(frames):
        out = oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), NC, 0, L, OFF, 0)
        for s, v in zip(state, out[:4]):
            s += v
    got = np.stack(_states(pq))
    assert np.array_equal(got[1:], state[1:])
    np.testing.assert_allclose(got[0], state[0], rtol=1e-14)
    assert np.array_equal(miou.confmat.cpu().numpy(), oracle.confmat((pred // L).numpy(), cat_t.numpy(), NC))

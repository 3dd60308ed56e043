"""Fused validation step: the kernel that writes the panoptic ids also feeds PQ + mIoU
(`PanopticPostprocessing.fuse_evaluation`, `npb_panoptic_forward_eval`).  Everything must be
identical to post-processing followed by `PanopticEvaluation.update`, and to the oracle."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

L, OFF = 1 << 16, 256 ** 3


def _make(C, dev, offset=OFF, with_mae=False):
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality,
                                                    PanopticQualityWithOrientationMAE)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori)()
    cls = PanopticQualityWithOrientationMAE if with_mae else PanopticQuality
    pq = cls(C + 1, 0, L, offset, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    return post, PanopticEvaluation(pq, miou), is_thing, has_ori


def _states(ev):
    return np.stack([getattr(ev.pq, n).cpu().numpy() for n in
                     ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])


def _raw(data, dev):
    d = {k: v.to(dev) for k, v in data.items()}
    return ((d['logits'], (d['heat'], d['offset'], d['orientation'])), (None, None))


@pytest.mark.parametrize('shape,offset', [((3, 96, 132), OFF),      # fused kernel
                                          ((2, 75, 91), OFF),       # H*W % 4 != 0: two passes
                                          ((2, 64, 80), 2 ** 25)])  # other offset: two passes
def test_fused_equals_separate_and_oracle(shape, offset, cuda_device):
    from nicr_mt_scene_analysis_b200 import testing
    B, H, W = shape
    C, K = 11, 5
    data = testing.make_batch(B, C, H, W, K, seed=B * H + W)
    post_f, ev_f, is_thing, has_ori = _make(C, cuda_device, offset)
    post_s, ev_s, _, _ = _make(C, cuda_device, offset)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')),
                                      is_thing, has_ori)
    tgt = np.roll(ref['panoptic'], 5, axis=-1)
    tgt_sem = (tgt // L).astype(np.uint8)
    batch = testing.make_batch_dict(B, H, W)
    gt = dict(batch, panoptic_fullres=torch.from_numpy(tgt).to(cuda_device),
              semantic_fullres=torch.from_numpy(tgt_sem).to(cuda_device))

    post_f.fuse_evaluation(ev_f)
    for _ in range(2):                                   # two batches accumulate
        r_f = post_f.postprocess(_raw(data, cuda_device), gt, is_training=False)
        r_s = post_s.postprocess(_raw(data, cuda_device), batch, is_training=False)
        assert r_f.get('_panoptic_evaluation_fused') is True
        assert '_panoptic_evaluation_fused' not in r_s
        ev_s.update(r_s['panoptic_segmentation_deeplab'], gt['panoptic_fullres'], gt['semantic_fullres'])
        for key in ('panoptic_segmentation_deeplab', 'panoptic_segmentation_deeplab_instance_idx',
                    'panoptic_segmentation_deeplab_semantic_idx'):
            assert torch.equal(r_f[key], r_s[key]), key
        assert r_f['panoptic_segmentation_deeplab_ids'] == r_s['panoptic_segmentation_deeplab_ids']
    assert np.array_equal(r_f['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
    ev_f.pq.check_status()
    ev_s.pq.check_status()
    assert np.array_equal(_states(ev_f), _states(ev_s))                  # bit-exact float64
    assert np.array_equal(ev_f.miou.confmat.cpu().numpy(), ev_s.miou.confmat.cpu().numpy())
    # ... and both equal the oracle (2 identical batches)
    state = np.zeros((4, C + 1))
    for _ in range(2):
        for b in range(B):
            out = oracle.pq_compare_and_accumulate(ref['panoptic'][b], tgt[b], C + 1, 0, L, offset, 0)
            for s, v in zip(state, out[:4]):
                s += v
    assert np.array_equal(_states(ev_f), state)
    assert np.array_equal(ev_f.miou.confmat.cpu().numpy(),
                          2 * oracle.confmat(ref['panoptic'] // L, tgt_sem, C + 1))
    assert int(ev_f.miou.confmat.sum()) == 2 * B * H * W


def test_fusion_needs_ground_truth_at_network_resolution(cuda_device):
    """No ground truth in the batch, or at another resolution: plain post-processing, the
    metric states stay untouched."""
    from nicr_mt_scene_analysis_b200 import testing
    B, C, H, W = 2, 7, 64, 80
    data = testing.make_batch(B, C, H, W, 4, seed=5)
    post, ev, _, _ = _make(C, cuda_device)
    post.fuse_evaluation(ev)
    batch = testing.make_batch_dict(B, H, W)
    r = post.postprocess(_raw(data, cuda_device), batch, is_training=False)
    assert '_panoptic_evaluation_fused' not in r
    other = dict(batch, panoptic_fullres=torch.zeros((B, 2 * H, 2 * W), dtype=torch.int64),
                 semantic_fullres=torch.zeros((B, 2 * H, 2 * W), dtype=torch.uint8))
    r = post.postprocess(_raw(data, cuda_device), other, is_training=False)
    assert '_panoptic_evaluation_fused' not in r
    assert float(ev.pq.tp_per_class.sum() + ev.pq.fn_per_class.sum() + ev.pq.fp_per_class.sum()) == 0.0
    assert int(ev.miou.confmat.sum()) == 0


def test_fused_validation_step_with_orientation(cuda_device):
    """Task helper + fused post-processing: PQ, mIoU and MAAE equal the unfused validation step."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.task_helper import PanopticTaskHelper
    B, C, H, W, K = 3, 9, 96, 128, 5
    data = testing.make_batch(B, C, H, W, K, seed=77)
    post_f, _, is_thing, has_ori = _make(C, cuda_device)
    post_s, _, _, _ = _make(C, cuda_device)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')),
                                      is_thing, has_ori)
    tgt = np.roll(ref['panoptic'], 3, axis=-1)
    tgt_ids = [{int(v): int(v) % L for v in np.unique(t) if int(v) % L} for t in tgt]
    ori_t = [{i: 0.3 * i - 1.0 for i in set(d.values())} for d in tgt_ids]
    batch = dict(testing.make_batch_dict(B, H, W),
                 panoptic_fullres=torch.from_numpy(tgt),
                 semantic_fullres=torch.from_numpy((tgt // L).astype(np.uint8)),
                 panoptic_ids_to_instance_dict=tgt_ids, orientations_present=ori_t)
    dev_batch = dict(batch, panoptic_fullres=batch['panoptic_fullres'].to(cuda_device),
                     semantic_fullres=batch['semantic_fullres'].to(cuda_device))
    helpers = []
    for post, fused in ((post_f, True), (post_s, False)):
        helper = PanopticTaskHelper(C + 1, (False,) + is_thing)
        helper.initialize(cuda_device)
        if fused:
            post.fuse_evaluation(helper.evaluation)
        r = post.postprocess(_raw(data, cuda_device), dev_batch, is_training=False)
        assert bool(r.get('_panoptic_evaluation_fused')) == fused
        assert ('_panoptic_matches' in r) == fused
        helper.validation_step(dev_batch, 1, r)
        helpers.append(helper.validation_epoch_end())
    (art_f, _, logs_f), (art_s, _, logs_s) = helpers
    assert set(logs_f) == set(logs_s) and set(art_f) == set(art_s)
    assert float(logs_f['panoptic_mae_deeplab_rad']) > 0
    for k in logs_s:
        if k.endswith('_time'):
            continue
        a, b = torch.as_tensor(logs_f[k]).double(), torch.as_tensor(logs_s[k]).double()
        if 'mae' in k:      # float64 sum of float32 errors in match order (not deterministic)
            assert torch.allclose(a, b, rtol=1e-12), k
        else:
            assert torch.equal(a, b), k
    for k in art_s:
        assert torch.equal(torch.as_tensor(art_f[k]).cpu(), torch.as_tensor(art_s[k]).cpu()) or \
            torch.allclose(torch.as_tensor(art_f[k]).cpu(), torch.as_tensor(art_s[k]).cpu(), equal_nan=True), k


def test_fused_step_with_overflowing_frames(cuda_device):
    """Ground truth of per-pixel random ids (far more than 4096 pairs per frame): the fused step
    defers those frames to the large-frame path, which reads the ids the step has written."""
    from nicr_mt_scene_analysis_b200 import testing
    B, C, H, W, K = 2, 11, 96, 128, 6
    data = testing.make_batch(B, C, H, W, K, seed=909)
    post, ev, is_thing, has_ori = _make(C, cuda_device)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')),
                                      is_thing, has_ori)
    g = torch.Generator().manual_seed(3)
    tgt_cat = torch.randint(1, C + 1, (B, H, W), generator=g)
    tgt = (tgt_cat * L + torch.randint(0, 60, (B, H, W), generator=g)).numpy()
    tgt_sem = tgt_cat.to(torch.uint8).numpy()
    gt = dict(testing.make_batch_dict(B, H, W), panoptic_fullres=torch.from_numpy(tgt).to(cuda_device),
              semantic_fullres=torch.from_numpy(tgt_sem).to(cuda_device))
    post.fuse_evaluation(ev)
    r = post.postprocess(_raw(data, cuda_device), gt, is_training=False)
    assert r['_panoptic_evaluation_fused']
    ev.pq.check_status()                                 # runs the large-frame path
    state = np.zeros((4, C + 1))
    n_pairs = 0
    for b in range(B):
        out = oracle.pq_compare_and_accumulate(ref['panoptic'][b], tgt[b], C + 1, 0, L, OFF, 0)
        n_pairs = max(n_pairs, len(np.unique(tgt[b].astype(np.int64) * OFF + ref['panoptic'][b])))
        for s, v in zip(state, out[:4]):
            s += v
    assert n_pairs > 4096                                # the case really overflows
    got = _states(ev)
    assert np.array_equal(got[1:], state[1:])
    np.testing.assert_allclose(got[0], state[0], rtol=1e-14)
    assert np.array_equal(ev.miou.confmat.cpu().numpy(), oracle.confmat(ref['panoptic'] // L, tgt_sem, C + 1))


def _gt_batch(pan_ref, B, H, W, dev, shift=5):
    from nicr_mt_scene_analysis_b200 import testing
    tgt = np.roll(pan_ref, shift, axis=-1)
    return dict(testing.make_batch_dict(B, H, W),
                panoptic_fullres=torch.from_numpy(tgt).to(dev),
                semantic_fullres=torch.from_numpy((tgt // L).astype(np.uint8)).to(dev))


def test_pipelined_matching_equals_in_place_matching(cuda_device):
    """`fuse_evaluation(..., pipeline_matching=True)`: the matcher of a batch runs during the next
    call (or when the states are read).  Batches of different sizes and contents, a read of the
    states in the middle, a stand-alone update in between and a reset: the float64 states and
    the confusion matrix are bit-identical to matching every batch in its own call."""
    from nicr_mt_scene_analysis_b200 import testing
    C, K, H, W = 9, 4, 64, 96
    post_p, ev_p, is_thing, has_ori = _make(C, cuda_device)
    post_s, ev_s, _, _ = _make(C, cuda_device)
    post_p.fuse_evaluation(ev_p, pipeline_matching=True)
    post_s.fuse_evaluation(ev_s)
    batches = []
    for i, B in enumerate((3, 3, 2, 3, 1, 3)):
        data = testing.make_batch(B, C, H, W, K, seed=60 + i)
        ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                            ('logits', 'heat', 'offset', 'orientation')),
                                          is_thing, has_ori)
        batches.append((data, _gt_batch(ref['panoptic'], B, H, W, cuda_device, shift=3 + i)))

    def run(post, ev, i):
        data, gt = batches[i]
        r = post.postprocess(_raw(data, cuda_device), gt, is_training=False)
        assert r.get('_panoptic_evaluation_fused') is True
        return r

    for i in range(3):
        rp, rs = run(post_p, ev_p, i), run(post_s, ev_s, i)
        assert torch.equal(rp['panoptic_segmentation_deeplab'], rs['panoptic_segmentation_deeplab'])
    assert ev_p.pq._deferred is not None            # the last batch is still unmatched ...
    assert np.array_equal(_states(ev_s), _states(ev_s))
    ev_p.pq.check_status()                          # ... until somebody needs the states
    assert ev_p.pq._deferred is None
    assert np.array_equal(_states(ev_p), _states(ev_s))
    # a stand-alone update between pipelined calls (it shares the hand-over workspace)
    run(post_p, ev_p, 3), run(post_s, ev_s, 3)
    data, gt = batches[0]
    pan = run(post_s, ev_s, 0)['panoptic_segmentation_deeplab']
    ev_p.update(pan, gt['panoptic_fullres'], gt['semantic_fullres'])
    for i in (4, 5):
        run(post_p, ev_p, i), run(post_s, ev_s, i)
    res_p, res_s = ev_p.compute(), ev_s.compute()
    assert np.array_equal(_states(ev_p), _states(ev_s))
    assert np.array_equal(ev_p.miou.confmat.cpu().numpy(), ev_s.miou.confmat.cpu().numpy())
    assert float(res_p['all_pq']) == float(res_s['all_pq'])
    # reset with a batch pending: that batch must not leak into the new states
    run(post_p, ev_p, 1)
    ev_p.reset()
    run(post_p, ev_p, 2)
    ev_s.reset()
    run(post_s, ev_s, 2)
    ev_p.pq.check_status(), ev_s.pq.check_status()
    assert np.array_equal(_states(ev_p), _states(ev_s))


def test_pipelined_matching_in_a_cuda_graph(cuda_device):
    """The pipelined step captured into a CUDA graph (what bench.py replays): every replay runs
    the matcher of the replay before it; reading the states flushes the last one; replaying after
    such a read, and eager calls mixed with replays, stay exact."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    C, K, B, H, W = 7, 3, 4, 48, 64
    post_p, ev_p, is_thing, has_ori = _make(C, cuda_device)
    post_s, ev_s, _, _ = _make(C, cuda_device)
    post_p.fuse_evaluation(ev_p, pipeline_matching=True)
    post_s.fuse_evaluation(ev_s)
    data = testing.make_batch(B, C, H, W, K, seed=71)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')), is_thing, has_ori)
    gt = _gt_batch(ref['panoptic'], B, H, W, cuda_device)
    raw = _raw(data, cuda_device)
    post_p._async_results = True                     # nothing may block inside a capture
    step = CapturedStep(lambda: post_p.postprocess(raw, gt, is_training=False), warmup=3,
                        device=cuda_device)
    n = 3                                            # the warm-up steps count as updates
    for _ in range(4):
        step.replay()
        n += 1
    ev_p.pq.check_status()
    for _ in range(2):                               # replays after a flush
        step.replay()
        n += 1
    post_p.postprocess(raw, gt, is_training=False)   # an eager call behind a replay
    n += 1
    step.replay()
    n += 1
    for _ in range(n):
        post_s.postprocess(raw, gt, is_training=False)
    res_p, res_s = ev_p.compute(), ev_s.compute()
    assert np.array_equal(_states(ev_p), _states(ev_s))
    assert np.array_equal(ev_p.miou.confmat.cpu().numpy(), ev_s.miou.confmat.cpu().numpy())
    assert float(res_p['all_pq']) == float(res_s['all_pq']) > 0


def test_pipelined_matching_keeps_orientation_batches_in_place(cuda_device):
    """batches that need their matches right away (orientation MAAE) are not deferred"""
    from nicr_mt_scene_analysis_b200 import testing
    C, K, B, H, W = 8, 3, 2, 48, 64
    post, ev, is_thing, has_ori = _make(C, cuda_device, with_mae=True)
    post.fuse_evaluation(ev, pipeline_matching=True)
    data = testing.make_batch(B, C, H, W, K, seed=81)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')), is_thing, has_ori)
    gt = _gt_batch(ref['panoptic'], B, H, W, cuda_device)
    post.postprocess(_raw(data, cuda_device), gt, is_training=False)
    assert ev.pq._deferred is not None
    gt_o = dict(gt, orientations_present=[{} for _ in range(B)])
    r = post.postprocess(_raw(data, cuda_device), gt_o, is_training=False)
    assert ev.pq._deferred is None and '_panoptic_matches' in r
    ev.pq.check_status()
    assert float(ev.pq.tp_per_class.sum()) > 0


@pytest.mark.parametrize('pipelined', [False, True])
def test_fused_evaluation_puts_large_frames_back_in_frame_order(pipelined, cuda_device, monkeypatch):
    """A validation loop through the fused (and pipelined) evaluation in which some frames have
    thousands of ground-truth segments (per-pixel ids in the upper rows: beyond the 1536
    segments / 4096 pairs of the batched matcher).  Such a frame is redone on the large-frame
    path and put back in its place: the float64 states equal the oracle's frame-by-frame sums
    (pq.py:298-303) bit for bit, the confusion matrix is untouched by the follow-up."""
    from nicr_mt_scene_analysis_b200 import testing
    C, K, H, W = 7, 4, 64, 96
    post, ev, is_thing, has_ori = _make(C, cuda_device)
    post.fuse_evaluation(ev, pipeline_matching=pipelined)
    from nicr_mt_scene_analysis_b200.metric import pq as pq_module
    redone = []
    real = pq_module._evaluate_big_frame
    monkeypatch.setattr(pq_module, '_evaluate_big_frame',
                        lambda *a, **k: (redone.append(1), real(*a, **k))[1])
    rng = np.random.default_rng(3)
    state = np.zeros((4, C + 1))
    cm = np.zeros((C + 1, C + 1), np.int64)
    noisy = {(1, 1), (3, 0), (3, 2), (4, 1)}        # (batch, frame) with a noisy target
    for i in range(6):
        B = 3
        data = testing.make_batch(B, C, H, W, K, seed=80 + i)
        ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                            ('logits', 'heat', 'offset', 'orientation')),
                                          is_thing, has_ori)
        tgt = np.roll(ref['panoptic'], 3 + i, axis=-1).copy()
        for b in range(B):
            if (i, b) in noisy:     # 40 rows of (nearly) unique thing ids of class 2
                tgt[b, :40] = 2 * L + 1 + rng.permutation(40 * W).reshape(40, W) % 3000
        tgt_sem = (tgt // L).astype(np.uint8)
        gt = dict(testing.make_batch_dict(B, H, W),
                  panoptic_fullres=torch.from_numpy(tgt).to(cuda_device),
                  semantic_fullres=torch.from_numpy(tgt_sem).to(cuda_device))
        r = post.postprocess(_raw(data, cuda_device), gt, is_training=False)
        assert r.get('_panoptic_evaluation_fused') is True
        for b in range(B):
            out = oracle.pq_compare_and_accumulate(ref['panoptic'][b], tgt[b], C + 1, 0, L, OFF, 0)
            for s, v in zip(state, out[:4]):
                s += v
        cm += oracle.confmat(ref['panoptic'] // L, tgt_sem, C + 1)
    ev.pq.check_status()
    assert len(redone) == len(noisy)                # every noisy frame took the large-frame path
    assert state[0].sum() != np.floor(state[0].sum())
    assert np.array_equal(_states(ev), state)
    assert np.array_equal(ev.miou.confmat.cpu().numpy(), cm)

"""Edge cases at the metric / post-processing boundary: empty inputs, a single pixel row, frames
without any thing pixel, the largest id geometry values the reference's helpers produce."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

L, OFF = 1 << 16, 256 ** 3


def test_empty_updates_add_nothing(cuda_device):
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticQuality
    m = MeanIntersectionOverUnion(n_classes=5, device=cuda_device)
    m.update(torch.empty(0, dtype=torch.int64, device=cuda_device),
             torch.empty(0, dtype=torch.uint8, device=cuda_device))
    m.update_nonvoid(torch.empty((0, 4, 4), dtype=torch.int64, device=cuda_device),
                     torch.empty((0, 4, 4), dtype=torch.uint8, device=cuda_device))
    m.check_status()
    assert int(m.confmat.sum()) == 0
    pq = PanopticQuality(3, 0, L, OFF, [False, True, False], device=cuda_device)
    pq.update(torch.empty((0, 8, 8), dtype=torch.int64, device=cuda_device),
              torch.empty((0, 8, 8), dtype=torch.int64, device=cuda_device))
    pq.check_status()
    for name in ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class'):
        assert float(getattr(pq, name).sum()) == 0.0


@pytest.mark.parametrize('shape', [(1, 1, 1), (1, 1, 7), (2, 3, 1), (1, 2, 130)])
def test_pq_and_miou_tiny_frames(shape, cuda_device):
    """Frames smaller than one load group (and widths that are not a multiple of 4)."""
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    NC = 4
    g = torch.Generator().manual_seed(sum(shape))
    cat_t = torch.randint(0, NC, shape, generator=g)
    cat_p = torch.randint(0, NC, shape, generator=g)
    is_thing = [False, True, False, True]
    it = torch.tensor(is_thing)
    tgt = cat_t * L + torch.where(it[cat_t], torch.randint(1, 3, shape, generator=g), torch.zeros(shape, dtype=torch.int64))
    pred = cat_p * L + torch.where(it[cat_p], torch.randint(1, 3, shape, generator=g), torch.zeros(shape, dtype=torch.int64))
    sem_t = cat_t.to(torch.uint8)
    pq = PanopticQuality(NC, 0, L, OFF, is_thing, device=cuda_device)
    miou = MeanIntersectionOverUnion(NC, True, device=cuda_device)
    PanopticEvaluation(pq, miou).update(pred.to(cuda_device), tgt.to(cuda_device), sem_t.to(cuda_device))
    pq.check_status()
    state = np.zeros((4, NC))
    for b in range(shape[0]):
        out = oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), NC, 0, L, OFF, 0)
        for s, v in zip(state, out[:4]):
            s += v
    got = np.stack([getattr(pq, n).cpu().numpy() for n in
                    ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])
    assert np.array_equal(got, state)
    assert np.array_equal(miou.confmat.cpu().numpy(), oracle.confmat(cat_p.numpy(), sem_t.numpy(), NC))


def test_postprocess_frame_without_things_and_void_only_target(cuda_device):
    """A batch whose semantic classes are all stuff (no foreground pixel at all, centres are
    still detected) next to a normal frame: ids = class * L everywhere, no instances, and the
    evaluation against an all-void target counts nothing but false positives of stuff."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    B, C, H, W, K = 2, 6, 48, 64, 3
    d = testing.make_batch(B, C, H, W, K, seed=11, with_orientation=False, device=cuda_device,
                           quantize='q10')
    is_thing = (False,) * C
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=(False,) * C)()
    r = post.postprocess(((d['logits'], (d['heat'], d['offset'])), (None, None)),
                         testing.make_batch_dict(B, H, W), is_training=False)
    ref = oracle.panoptic_postprocess(d['logits'].cpu().numpy(), d['heat'].cpu().numpy(),
                                      d['offset'].cpu().numpy(), None, is_thing, (False,) * C)
    pan = r['panoptic_segmentation_deeplab'].cpu().numpy()
    assert np.array_equal(pan, ref['panoptic'])
    assert np.array_equal(pan % L, np.zeros_like(pan))
    assert int(r['panoptic_segmentation_deeplab_instance_idx'].sum()) == 0
    assert r['panoptic_segmentation_deeplab_ids'] == [{}, {}]
    assert not bool(r['panoptic_foreground_mask'].any())
    pq = PanopticQuality(C + 1, 0, L, OFF, (False,) + is_thing, device=cuda_device)
    pq.update(r['panoptic_segmentation_deeplab'].to(cuda_device),
              torch.zeros((B, H, W), dtype=torch.int64, device=cuda_device))
    pq.check_status()
    # every predicted segment lies completely in the ignored (void) ground truth: not a FP
    assert float(pq.fp_per_class.sum()) == 0.0 and float(pq.tp_per_class.sum()) == 0.0
    assert float(pq.fn_per_class.sum()) == 0.0
    for b in range(B):
        out = oracle.pq_compare_and_accumulate(pan[b], np.zeros((H, W), np.int64), C + 1, 0, L, OFF, 0)
        assert all(float(v.sum()) == 0.0 for v in out[:4])


@pytest.mark.parametrize('hw', [(75, 91), (64, 96)])      # odd map size: scalar kernels; 4 px / thread
def test_score_maps_against_a_torch_restatement(hw, cuda_device):
    """compute_scores=True (panoptic.py:171-239): semantic score = soft-max probability of the
    pixel's panoptic class (0 for void), instance score = heat-map value at the instance's centre,
    panoptic score = mean semantic score of the instance x instance score (things) or the
    semantic score (stuff); restated with torch fp32 on the kernel's own label maps, 1e-5."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    H, W = hw
    B, C, K = 2, 7, 4
    d = testing.make_batch(B, C, H, W, K, seed=23, with_orientation=False, device=cuda_device,
                           quantize='q10')
    is_thing = testing.default_is_thing(C)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=(False,) * C,
        compute_scores=True)()
    r = post.postprocess(((d['logits'], (d['heat'], d['offset'])), (None, None)),
                         testing.make_batch_dict(B, H, W), is_training=False)
    pan_sem = r['panoptic_segmentation_deeplab_semantic_idx'].cpu()           # 0 = void
    inst = r['panoptic_segmentation_deeplab_instance_idx'].cpu().long()
    probs = torch.softmax(d['logits'].cpu(), dim=1)
    want = (pan_sem - 1).clamp(min=0)
    sem_score = torch.gather(probs, 1, want[:, None])[:, 0] * (pan_sem > 0)
    np.testing.assert_allclose(r['panoptic_segmentation_deeplab_semantic_score'].cpu().numpy(),
                               sem_score.numpy(), rtol=1e-5, atol=1e-7)
    pan_score = sem_score.clone()
    inst_score = torch.zeros_like(sem_score)
    meta = r['panoptic_segmentation_deeplab_instance_meta']
    ids = r['panoptic_segmentation_deeplab_ids']
    for b in range(B):
        kept = set(ids[b].values())                  # instances that survived the merge
        for i, m in meta[b].items():
            mask = inst[b] == i
            if not bool(mask.any()):
                continue
            if i in kept:       # instances dropped by the merge keep 0 / the semantic score
                inst_score[b][mask] = m['score']
                pan_score[b][mask] = sem_score[b][mask].double().mean().float() * m['score']
    np.testing.assert_allclose(r['panoptic_segmentation_deeplab_instance_score'].cpu().numpy(),
                               inst_score.numpy(), rtol=1e-6)
    np.testing.assert_allclose(r['panoptic_segmentation_deeplab_panoptic_score'].cpu().numpy(),
                               pan_score.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('C,hw', [(1, (5, 7)), (19, (33, 40)), (40, (31, 37)), (64, (8, 12)), (70, (16, 20)),
                                  (70, (15, 21))])
def test_softmax_scores_all_kernels(C, hw, cuda_device):
    """`semantic_softmax_scores` (two passes, 4 px / thread and the scalar form for odd map sizes)
    against torch.softmax, 1e-5 relative; -inf and NaN logits behave like the reference's."""
    from nicr_mt_scene_analysis_b200.model.postprocessing.semantic import softmax_scores
    g = torch.Generator().manual_seed(C)
    logits = (torch.randn((2, C) + hw, generator=g) * 4).to(cuda_device)
    if C > 1:
        logits[0, 0, 0, 0] = float('-inf')
    got = softmax_scores(logits)
    ref = torch.softmax(logits.cpu(), dim=1)
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(got.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    if C > 1:       # a NaN logit poisons its column like the reference's soft-max
        logits[1, 1, 2, 3] = float('nan')
        got = softmax_scores(logits)
        assert bool(torch.isnan(got[1, :, 2, 3]).all())
        assert not bool(torch.isnan(got[1, :, 2, 4]).any())


def test_tensors_on_a_device_that_is_not_current(cuda_device):
    """One process, several GPUs: post-processing and metrics of tensors on cuda:1 while cuda:0 is
    the current device give the results of cuda:0 and leave the current device alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    B, C, H, W, K = 2, 6, 48, 64, 3
    d = testing.make_batch(B, C, H, W, K, seed=5, with_orientation=True, quantize='q10')
    is_thing = testing.default_is_thing(C)
    results = []
    torch.cuda.set_device(0)
    for index in (0, 1):
        dev = torch.device('cuda', index)
        post = get_postprocessing_class(
            'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
            instance_postprocessing=get_postprocessing_class('instance')(),
            semantic_classes_is_thing=is_thing, semantic_class_has_orientation=is_thing)()
        r = post.postprocess(((d['logits'].to(dev), (d['heat'].to(dev), d['offset'].to(dev),
                                                     d['orientation'].to(dev))), (None, None)),
                             testing.make_batch_dict(B, H, W), is_training=False)
        pan = r['panoptic_segmentation_deeplab']
        assert pan.device == dev and torch.cuda.current_device() == 0
        tgt = torch.roll(pan, 3, -1).contiguous()
        pq = PanopticQuality(C + 1, 0, L, OFF, (False,) + is_thing, device=dev)
        miou = MeanIntersectionOverUnion(C + 1, True, device=dev)
        PanopticEvaluation(pq, miou).update(pan, tgt, (tgt // L).to(torch.uint8))
        res = pq.compute()
        assert torch.cuda.current_device() == 0
        results.append((pan.cpu(), r['panoptic_segmentation_deeplab_ids'], miou.confmat.cpu(),
                        res['all_pq'], pq.iou_per_class.cpu()))
    for a, b in zip(*results):
        if isinstance(a, torch.Tensor):
            assert torch.equal(a, b)
        else:
            assert a == b


def test_memory_formats_and_misaligned_views(cuda_device):
    """channels_last decoder outputs (NCHW shape, NHWC strides) and contiguous views whose storage
    starts 4 bytes off a 16-byte boundary give the results of plain contiguous inputs (the
    kernels fall back to their scalar forms / the host makes one contiguous copy)."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    B, C, H, W, K = 2, 8, 48, 64, 4
    d = testing.make_batch(B, C, H, W, K, seed=9, with_orientation=True, device=cuda_device,
                           quantize='q10')
    is_thing = testing.default_is_thing(C)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=is_thing)()

    def run(t):
        r = post.postprocess(((t['logits'], (t['heat'], t['offset'], t['orientation'])), (None, None)),
                             testing.make_batch_dict(B, H, W), is_training=False)
        return (r['panoptic_segmentation_deeplab'].clone(), r['semantic_segmentation_idx'].clone(),
                r['panoptic_segmentation_deeplab_instance_idx'].clone(),
                r['panoptic_segmentation_deeplab_ids'],
                r['orientations_panoptic_segmentation_deeplab_instance'])

    want = run(d)

    def shifted(x):         # same values, contiguous, storage offset of one element (4 bytes)
        flat = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
        flat[1:].copy_(x.reshape(-1))
        v = flat[1:].view(x.shape)
        assert v.is_contiguous() and v.data_ptr() % 16 == 4
        return v

    variants = {
        'channels_last': {k: v.contiguous(memory_format=torch.channels_last) for k, v in d.items()},
        'shifted': {k: shifted(v) for k, v in d.items()},
        'shifted logits only': dict(d, logits=shifted(d['logits'])),
        'shifted offset only': dict(d, offset=shifted(d['offset'])),
    }
    for name, t in variants.items():
        got = run(t)
        for a, b in zip(want[:3], got[:3]):
            assert torch.equal(a, b), name
        assert want[3] == got[3], name
        for dw, dg in zip(want[4], got[4]):
            assert sorted(dw) == sorted(dg), name
            for k in dw:
                assert abs(dw[k] - dg[k]) <= 1e-5 * max(1.0, abs(dw[k])), name

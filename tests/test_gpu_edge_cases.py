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

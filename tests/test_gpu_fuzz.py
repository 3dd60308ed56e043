"""Seeded random sweep: the whole CUDA path (post-processing + evaluation) against the oracle
on many small, odd-shaped, tie-heavy configurations.  Everything integer bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

N_SEEDS = int(os.environ.get('NPB_FUZZ_SEEDS', '40'))      # a longer sweep on request


def _cfg(rng):
    H = int(rng.integers(12, 90))
    W = int(rng.integers(12, 110))
    return dict(
        B=int(rng.integers(1, 4)), C=int(rng.integers(1, 24)), H=H, W=W,
        K=int(rng.integers(0, 12)), quantize=str(rng.choice(['q10', 'tie'])),
        top_k=int(rng.integers(1, 12)), ks=int(rng.choice([1, 3, 3, 3, 5, 7])),
        thr=float(rng.choice([0.1, 0.1, 0.3, 0.6, -0.5])), apply_fg=bool(rng.integers(0, 2)),
        normalized=bool(rng.integers(0, 2)),
        dist_thr=(None if rng.integers(0, 2) else int(rng.integers(2, 30))),
        with_orientation=bool(rng.integers(0, 2)))


@pytest.mark.parametrize('seed', range(N_SEEDS))
def test_random_configuration(seed, cuda_device):
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    rng = np.random.default_rng(1000 + seed)
    c = _cfg(rng)
    B, C, H, W = c['B'], c['C'], c['H'], c['W']
    if min(H, W) <= 16 + 1:
        c['K'] = 0                       # make_frame keeps centres 8 px off the border
    data = testing.make_batch(B, C, H, W, max(c['K'], 1), seed=seed, quantize=c['quantize'],
                              with_orientation=c['with_orientation'])
    if c['K'] == 0:
        data['heat'].zero_()
    if not c['normalized']:
        data['offset'][:, 0] *= H
        data['offset'][:, 1] *= W
    is_thing = tuple(bool(x) for x in rng.integers(0, 2, C))
    has_ori = tuple(bool(t and rng.integers(0, 2)) for t in is_thing)
    pan = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class(
            'instance', heatmap_threshold=c['thr'], heatmap_nms_kernel_size=c['ks'],
            top_k_instances=c['top_k'], heatmap_apply_foreground_mask=c['apply_fg'],
            normalized_offset=c['normalized'], offset_distance_threshold=c['dist_thr'])(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
        normalized_offset=c['normalized'])()
    dev = cuda_device
    inst_out = (data['heat'].to(dev), data['offset'].to(dev)) + \
        ((data['orientation'].to(dev),) if c['with_orientation'] else ())
    ref_kwargs = dict(threshold=c['thr'], nms_kernel_size=c['ks'], top_k=c['top_k'],
                      apply_foreground_mask=c['apply_fg'], normalized_offset=c['normalized'],
                      offset_distance_threshold=c['dist_thr'])
    try:
        ref = oracle.panoptic_postprocess(
            data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(),
            data['orientation'].numpy() if c['with_orientation'] else None, is_thing, has_ori,
            **ref_kwargs)
    except oracle.OracleError as e:
        assert e.code == -2              # > 255 centres: both sides must refuse
        from nicr_mt_scene_analysis_b200 import _lib
        with pytest.raises(_lib.NpbError):
            pan.postprocess(((data['logits'].to(dev), inst_out), (None, None)),
                            testing.make_batch_dict(B, H, W), is_training=False)
        return
    r = pan.postprocess(((data['logits'].to(dev), inst_out), (None, None)),
                        testing.make_batch_dict(B, H, W), is_training=False)
    assert np.array_equal(r['_semantic_segmentation_idx_u8'].cpu().numpy(), ref['semantic_idx']), c
    assert np.array_equal(r['panoptic_segmentation_deeplab_instance_idx'].cpu().numpy(),
                          ref['instance_idx']), c
    assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic']), c
    assert r['panoptic_segmentation_deeplab_ids'] == ref['ids'], c
    for gm, rm in zip(r['panoptic_segmentation_deeplab_instance_meta'], ref['meta']):
        assert {k: (v['center_yx'], v['area']) for k, v in gm.items()} == \
            {k: (v['center_yx'], v['area']) for k, v in rm.items()}, c
    if c['with_orientation']:
        for dg, dr in zip(r['orientations_panoptic_segmentation_deeplab_instance'],
                          ref['orientations']):
            assert sorted(dg) == sorted(dr), c
            for k in dr:
                assert abs(dg[k] - dr[k]) <= 1e-5 * max(1.0, abs(dr[k])), (c, k)

    # evaluation of the prediction against a shifted copy, standard and odd id geometries
    pred = r['panoptic_segmentation_deeplab']
    tgt = torch.roll(pred, int(rng.integers(1, 6)), dims=-1).contiguous()
    tgt_sem = (tgt // 65536).to(torch.uint8)
    pq = PanopticQuality(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    PanopticEvaluation(pq, miou).update(pred, tgt, tgt_sem)
    pq2 = PanopticQuality(C + 1, int(rng.integers(0, C + 1)), 1 << 16, 3 * 10 ** 7,
                          (False,) + is_thing, device=dev)        # offset not a power of two
    pq2.update(pred, tgt)
    # the same batch through the fused step (ids written and evaluated by one kernel)
    pq3 = PanopticQuality(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev)
    miou3 = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    pan.fuse_evaluation(PanopticEvaluation(pq3, miou3))
    rf = pan.postprocess(((data['logits'].to(dev), inst_out), (None, None)),
                         dict(testing.make_batch_dict(B, H, W), panoptic_fullres=tgt,
                              semantic_fullres=tgt_sem), is_training=False)
    assert rf['_panoptic_evaluation_fused']
    assert torch.equal(rf['panoptic_segmentation_deeplab'], pred), c
    assert torch.equal(rf['panoptic_segmentation_deeplab_semantic_idx'],
                       r['panoptic_segmentation_deeplab_semantic_idx']), c
    assert torch.equal(miou3.confmat, miou.confmat), c
    # the stand-alone confusion-matrix kernels on the same maps (int64 predictions, uint8 targets):
    # all pixels, and the semantic task helper's form (void skipped, target - 1)
    miou4 = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    miou4.update(r['panoptic_segmentation_deeplab_semantic_idx'], tgt_sem)
    miou4.check_status()
    assert torch.equal(miou4.confmat, miou.confmat), c
    miou5 = MeanIntersectionOverUnion(C, device=dev)
    miou5.update_nonvoid(r['semantic_segmentation_idx'], tgt_sem)
    miou5.check_status()
    keep = (tgt_sem != 0).cpu().numpy()
    assert np.array_equal(miou5.confmat.cpu().numpy(),
                          oracle.confmat(ref['semantic_idx'][keep], tgt_sem.cpu().numpy()[keep].astype(np.int64) - 1, C)), c
    for metric in (pq, pq2, pq3):
        state = np.zeros((4, C + 1))
        zero_division = False
        for b in range(B):
            try:
                out = oracle.pq_compare_and_accumulate(
                    ref['panoptic'][b], tgt[b].cpu().numpy(), C + 1, metric.ignored_label, 1 << 16,
                    metric.offset, metric.void_segment_id)
            except oracle.OracleError as e:
                assert e.code == -3          # union == 0: the reference raises ZeroDivisionError
                zero_division = True
                break
            for s, v in zip(state, out[:4]):
                s += v
        if zero_division:
            with pytest.raises(ZeroDivisionError):
                metric.check_status()
            continue
        metric.check_status()
        got = np.stack([getattr(metric, n).cpu().numpy() for n in
                        ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])
        assert np.array_equal(got, state), c
    cm = oracle.confmat(ref['panoptic'] // 65536, tgt_sem.cpu().numpy(), C + 1)
    assert np.array_equal(miou.confmat.cpu().numpy(), cm), c

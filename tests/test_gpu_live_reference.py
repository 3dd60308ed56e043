"""The CUDA path against the UNMODIFIED reference, run live on the GPU box.

`baseline/_ref/` (the reference as `pip install --target` leaves it, staged by
`__graft_entry__.build()` in the authoring container, git-ignored, shipped with the gpurun
snapshot) is imported through the stand-ins of oracle/ref_stubs for its two absent third-party
dependencies.  No oracle in between: post-processing results, PQ states (float64, bit for bit)
and confusion matrices of the product are compared with what the reference computes on the host
for the same random decoder outputs."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import reference_arm  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_arm.available(), reason='reference not staged')]

L, OFF = 1 << 16, 256 ** 3


@pytest.fixture(scope='module')
def ref():
    threads = torch.get_num_threads()
    torch.set_num_threads(min(8, len(os.sched_getaffinity(0))))
    yield reference_arm.load()
    torch.set_num_threads(threads)


def _cfg(rng):
    return dict(
        B=int(rng.integers(1, 4)), C=int(rng.integers(2, 41)), H=int(rng.integers(24, 100)),
        W=int(rng.integers(24, 140)), K=int(rng.integers(0, 9)),
        quantize=str(rng.choice(['q10', 'tie', 'q10'])), top_k=int(rng.integers(1, 12)),
        ks=int(rng.choice([1, 3, 3, 5, 7])), thr=float(rng.choice([0.1, 0.3, 0.6, -0.5])),
        apply_fg=bool(rng.integers(0, 2)), normalized=bool(rng.integers(0, 2)),
        dist_thr=(None if rng.integers(0, 2) else int(rng.integers(3, 25))),
        with_orientation=bool(rng.integers(0, 2)), poison=bool(rng.integers(0, 4) == 0))


@pytest.mark.parametrize('seed', range(int(os.environ.get('NPB_LIVE_SEEDS', '0')) or 12))
def test_cuda_path_equals_live_reference(seed, ref, cuda_device):
    from nicr_mt_scene_analysis_b200 import _lib, testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    rng = np.random.default_rng(4200 + seed)
    c = _cfg(rng)
    B, C, H, W = c['B'], c['C'], c['H'], c['W']
    data = testing.make_batch(B, C, H, W, max(c['K'], 1), seed=900 + seed, quantize=c['quantize'],
                              with_orientation=c['with_orientation'])
    if c['K'] == 0:
        data['heat'].zero_()
    if not c['normalized']:
        data['offset'][:, 0] *= H
        data['offset'][:, 1] *= W
    if c['poison']:
        testing.poison_logits(data['logits'], 0.05, seed)
    is_thing = tuple(bool(x) for x in rng.integers(0, 2, C))
    has_ori = tuple(bool(t and rng.integers(0, 2)) for t in is_thing)
    kw = dict(heatmap_threshold=c['thr'], heatmap_nms_kernel_size=c['ks'], top_k_instances=c['top_k'],
              heatmap_apply_foreground_mask=c['apply_fg'], normalized_offset=c['normalized'],
              offset_distance_threshold=c['dist_thr'])

    def build(get):
        return get('panoptic', semantic_postprocessing=get('semantic')(),
                   instance_postprocessing=get('instance', **kw)(),
                   semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
                   normalized_offset=c['normalized'])()

    names = ('heat', 'offset') + (('orientation',) if c['with_orientation'] else ())
    batch = testing.make_batch_dict(B, H, W)
    # ---- the reference, on the host
    want = build(ref['get_postprocessing_class']).postprocess(
        ((data['logits'].clone(), tuple(data[k].clone() for k in names)), (None, None)), batch,
        is_training=False)
    # ---- the product, on the GPU
    raw = ((data['logits'].to(cuda_device), tuple(data[k].to(cuda_device) for k in names)),
           (None, None))
    try:
        got = build(get_postprocessing_class).postprocess(raw, batch, is_training=False)
    except _lib.NpbError as e:
        # the one deliberate deviation on this path: > 255 centres in a frame are refused, the
        # reference wraps its uint8 instance ids silently (instance.py:236)
        assert e.code == _lib.ERR_TOO_MANY_CENTERS
        assert max(len(m) for m in want['panoptic_segmentation_deeplab_instance_meta']) > 255, c
        return
    for key in ('semantic_segmentation_idx', 'panoptic_segmentation_deeplab',
                'panoptic_segmentation_deeplab_instance_idx',
                'panoptic_segmentation_deeplab_semantic_idx', 'panoptic_foreground_mask'):
        assert torch.equal(got[key].cpu(), want[key].cpu()), (key, c)
        assert got[key].dtype == want[key].dtype, (key, c)
    assert got['panoptic_segmentation_deeplab_ids'] == \
        [{int(k): int(v) for k, v in d.items()} for d in want['panoptic_segmentation_deeplab_ids']], c
    for gm, wm in zip(got['panoptic_segmentation_deeplab_instance_meta'],
                      want['panoptic_segmentation_deeplab_instance_meta']):
        assert sorted(gm) == sorted(int(k) for k in wm), c
        for k, v in wm.items():
            g = gm[int(k)]
            assert tuple(g['center_yx']) == tuple(int(x) for x in v['center_yx']), (c, k)
            assert g['area'] == int(v['area']), (c, k)
            assert g['score'] == pytest.approx(float(v['score']), rel=1e-6), (c, k)
    np.testing.assert_allclose(got['semantic_segmentation_score'].cpu().numpy(),
                               want['semantic_segmentation_score'].numpy(), rtol=1e-5)
    if c['with_orientation']:
        key = 'orientations_panoptic_segmentation_deeplab_instance'
        for dg, dw in zip(got[key], want[key]):
            assert sorted(dg) == sorted(int(k) for k in dw), c
            for k, v in dw.items():         # north_star: 1e-5 relative
                assert abs(dg[int(k)] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (c, k)

    # ---- evaluation: PanopticQuality / mIoU objects of both sides fed with the same maps
    pred = want['panoptic_segmentation_deeplab']
    tgt = torch.roll(pred, int(rng.integers(1, 6)), dims=-1).contiguous()
    tgt_sem = (tgt // L).to(torch.uint8)
    thing = [False] + list(is_thing)
    ref_state = [torch.zeros(C + 1, dtype=torch.float64) for _ in range(4)]
    try:
        for b in range(B):      # the reference's per-frame worker, in frame order (pq.py:298-303)
            out = ref['compare_and_accumulate'](pred[b], tgt[b], C + 1, 0, L, OFF, 0)
            for s, v in zip(ref_state, out[:4]):
                s += v
    except ZeroDivisionError:
        pq = PanopticQuality(C + 1, 0, L, OFF, thing, device=cuda_device)
        pq.update(got['panoptic_segmentation_deeplab'], tgt.to(cuda_device))
        with pytest.raises(ZeroDivisionError):
            pq.check_status()
        return
    ref_miou = ref['MeanIntersectionOverUnion'](n_classes=C + 1, ignore_first_class=True)
    ref_miou.reset()
    ref_miou.update(preds=pred // L, target=tgt_sem)
    pq = PanopticQuality(C + 1, 0, L, OFF, thing, device=cuda_device)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=cuda_device)
    PanopticEvaluation(pq, miou).update(got['panoptic_segmentation_deeplab'], tgt.to(cuda_device),
                                        tgt_sem.to(cuda_device))
    pq.check_status()
    for name, w_ in zip(('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class'), ref_state):
        assert np.array_equal(getattr(pq, name).cpu().numpy(), w_.numpy()), (name, c)   # bit for bit
    assert np.array_equal(miou.confmat.cpu().numpy(), ref_miou.confmat.numpy()), c
    assert float(miou.compute()) == pytest.approx(float(ref_miou.compute()), rel=1e-6)


@pytest.mark.parametrize('seed', range(int(os.environ.get('NPB_LIVE_SEEDS', '0')) or 6))
def test_wrapped_instance_ids_equal_live_reference(seed, ref, cuda_device):
    """`on_overflow='wrap'` against the unmodified reference in the regime where its uint8
    instance ids wrap (instance.py:231-236): lattices of exactly tied heat-map peaks, hundreds of
    centres per frame (one frame of the batch stays ordinary), random NMS windows, foreground-
    masked centres, distance thresholds, offset scales."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    rng = np.random.default_rng(5300 + seed)
    B, C = 3, int(rng.integers(3, 12))
    H, W = int(rng.integers(66, 120)), int(rng.integers(70, 150))
    ks = int(rng.choice([1, 3, 3, 5]))
    step = int(rng.choice([3, 4])) if ks <= 3 else 4
    apply_fg = bool(rng.integers(0, 2))
    dist_thr = None if rng.integers(0, 2) else int(rng.integers(3, 40))
    normalized = bool(rng.integers(0, 2))
    data = testing.make_batch(B, C, H, W, 4, seed=700 + seed, quantize='q10', with_orientation=True)
    testing.saturate_heat(data['heat'], step=step, value=float(rng.choice([1.0, 0.5])), frames=(0, 2))
    data['offset'][2] *= float(rng.choice([0.0, 0.01, 0.3]))
    if not normalized:
        data['offset'][:, 0] *= H
        data['offset'][:, 1] *= W
    is_thing = (True,) * C if apply_fg else tuple(bool(x) for x in rng.integers(0, 2, C))
    if not any(is_thing):
        is_thing = (True,) + is_thing[1:]
    has_ori = tuple(bool(t and rng.integers(0, 2)) for t in is_thing)
    kw = dict(heatmap_nms_kernel_size=ks, top_k_instances=int(rng.integers(1, 65)),
              heatmap_apply_foreground_mask=apply_fg, normalized_offset=normalized,
              offset_distance_threshold=dist_thr)

    def build(get, **extra):
        return get('panoptic', semantic_postprocessing=get('semantic')(),
                   instance_postprocessing=get('instance', **kw, **extra)(),
                   semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
                   normalized_offset=normalized)()

    names = ('heat', 'offset', 'orientation')
    batch = testing.make_batch_dict(B, H, W)
    want = build(ref['get_postprocessing_class']).postprocess(
        ((data['logits'].clone(), tuple(data[k].clone() for k in names)), (None, None)), batch,
        is_training=False)
    wmeta = want['panoptic_segmentation_deeplab_instance_meta']
    assert max(len(m) for m in wmeta) > 255, [len(m) for m in wmeta]
    raw = ((data['logits'].to(cuda_device), tuple(data[k].to(cuda_device) for k in names)),
           (None, None))
    got = build(get_postprocessing_class, on_overflow='wrap').postprocess(raw, batch,
                                                                         is_training=False)
    for key in ('semantic_segmentation_idx', 'panoptic_segmentation_deeplab',
                'panoptic_segmentation_deeplab_instance_idx',
                'panoptic_segmentation_deeplab_semantic_idx'):
        assert torch.equal(got[key].cpu(), want[key].cpu()), key
    assert got['panoptic_segmentation_deeplab_ids'] == \
        [{int(k): int(v) for k, v in d.items()} for d in want['panoptic_segmentation_deeplab_ids']]
    for gm, wm in zip(got['panoptic_segmentation_deeplab_instance_meta'], wmeta):
        assert sorted(gm) == sorted(int(k) for k in wm)
        for k, v in wm.items():
            g = gm[int(k)]
            assert tuple(g['center_yx']) == tuple(int(x) for x in v['center_yx']), k
            assert g['area'] == int(v['area']), k
            assert g['score'] == pytest.approx(float(v['score']), rel=1e-6), k
            wo, go = float(v['orientation']), g['orientation']
            assert (wo != wo and go != go) or abs(go - wo) <= 1e-5 * max(1.0, abs(wo)), k
    key = 'orientations_panoptic_segmentation_deeplab_instance'
    for dg, dw in zip(got[key], want[key]):
        assert sorted(dg) == sorted(int(k) for k in dw)
        for k, v in dw.items():
            assert abs(dg[int(k)] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), k

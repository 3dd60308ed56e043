"""The oracle against the UNMODIFIED reference, run live on fresh random configurations.

Only where /root/reference exists (the authoring container; the GPU box has no reference and
skips this file): the reference is imported through the stub packages in oracle/ref_stubs
(stand-ins for the absent torchmetrics / nicr_scene_analysis_datasets, no arithmetic).  This
pins the oracle beyond the committed golden vectors: ids, maps, id dicts, centres, areas,
PQ states (float64, bit for bit) and confusion matrices must be identical."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference/src'
MORE = int(os.environ.get('NPB_LIVE_SEEDS', '0'))      # a longer sweep on request

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason='reference sources not present')


@pytest.fixture(scope='module')
def ref():
    """The reference's entry points (imported once; appended to sys.path so that nothing of the
    test environment is shadowed by the stubs)."""
    for p in (os.path.join(ROOT, 'oracle', 'ref_stubs'), REF_SRC):
        if p not in sys.path:
            sys.path.append(p)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    from nicr_mt_scene_analysis.metric import MeanIntersectionOverUnion
    from nicr_mt_scene_analysis.metric.pq import compare_and_accumulate
    from nicr_mt_scene_analysis.model.postprocessing import get_postprocessing_class
    yield dict(get=get_postprocessing_class, pq=compare_and_accumulate, miou=MeanIntersectionOverUnion)
    torch.set_num_threads(threads)


def _cfg(rng):
    return dict(
        B=int(rng.integers(1, 3)), C=int(rng.integers(2, 12)), H=int(rng.integers(24, 72)),
        W=int(rng.integers(24, 90)), K=int(rng.integers(0, 7)),
        quantize=str(rng.choice(['q10', 'tie'])), top_k=int(rng.integers(1, 9)),
        ks=int(rng.choice([1, 3, 3, 5, 7])), thr=float(rng.choice([0.1, 0.3, 0.6, -0.5])),
        apply_fg=bool(rng.integers(0, 2)), normalized=bool(rng.integers(0, 2)),
        dist_thr=(None if rng.integers(0, 2) else int(rng.integers(3, 25))),
        with_orientation=bool(rng.integers(0, 2)))


@pytest.mark.parametrize('seed', range(MORE or 8))
def test_oracle_equals_live_reference(seed, ref):
    from nicr_mt_scene_analysis_b200 import testing
    rng = np.random.default_rng(7000 + seed)
    c = _cfg(rng)
    B, C, H, W = c['B'], c['C'], c['H'], c['W']
    data = testing.make_batch(B, C, H, W, max(c['K'], 1), seed=300 + seed, quantize=c['quantize'],
                              with_orientation=c['with_orientation'])
    if c['K'] == 0:
        data['heat'].zero_()
    if not c['normalized']:
        data['offset'][:, 0] *= H
        data['offset'][:, 1] *= W
    is_thing = tuple(bool(x) for x in rng.integers(0, 2, C))
    has_ori = tuple(bool(t and rng.integers(0, 2)) for t in is_thing)
    get = ref['get']
    pan = get('panoptic', semantic_postprocessing=get('semantic')(),
              instance_postprocessing=get(
                  'instance', heatmap_threshold=c['thr'], heatmap_nms_kernel_size=c['ks'],
                  top_k_instances=c['top_k'], heatmap_apply_foreground_mask=c['apply_fg'],
                  normalized_offset=c['normalized'], offset_distance_threshold=c['dist_thr'])(),
              semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
              normalized_offset=c['normalized'])()
    inst_out = (data['heat'], data['offset']) + ((data['orientation'],) if c['with_orientation'] else ())
    r = pan.postprocess(((data['logits'].clone(), tuple(t.clone() for t in inst_out)), (None, None)),
                        testing.make_batch_dict(B, H, W), is_training=False)
    try:
        got = oracle.panoptic_postprocess(
            data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(),
            data['orientation'].numpy() if c['with_orientation'] else None, is_thing, has_ori,
            threshold=c['thr'], nms_kernel_size=c['ks'], top_k=c['top_k'],
            apply_foreground_mask=c['apply_fg'], normalized_offset=c['normalized'],
            offset_distance_threshold=c['dist_thr'])
    except oracle.OracleError as e:
        # the one deliberate deviation on this path: more than 255 centres
        # in a frame are refused, the reference wraps its uint8 ids silently (instance.py:236)
        assert e.code == -2
        _, centers = pan._instance_postprocessing._get_instance_centers(
            data['heat'], r['panoptic_foreground_mask'])
        assert max(len(x) for x in centers) > 255, c
        return
    assert np.array_equal(got['semantic_idx'], r['semantic_segmentation_idx'].numpy()), c
    assert np.array_equal(got['instance_idx'], r['panoptic_segmentation_deeplab_instance_idx'].numpy()), c
    assert np.array_equal(got['panoptic'], r['panoptic_segmentation_deeplab'].numpy()), c
    assert got['ids'] == [{int(k): int(v) for k, v in d.items()}
                          for d in r['panoptic_segmentation_deeplab_ids']], c
    for gm, rm in zip(got['meta'], r['panoptic_segmentation_deeplab_instance_meta']):
        assert {k: (tuple(v['center_yx']), v['area']) for k, v in gm.items()} == \
            {int(k): (tuple(int(x) for x in v['center_yx']), int(v['area'])) for k, v in rm.items()}, c
    for gm, rm in zip(got['meta'], r['panoptic_segmentation_deeplab_instance_meta']):
        for k, v in rm.items():             # heat-map value at the centre (instance.py:262)
            assert gm[int(k)]['score'] == pytest.approx(float(v['score']), rel=1e-6), (c, k)
    np.testing.assert_allclose(oracle.semantic_score(data['logits'].numpy()),
                               r['semantic_segmentation_score'].numpy(), rtol=1e-5)
    if c['with_orientation']:
        for dg, dr in zip(got['orientations'], r['orientations_panoptic_segmentation_deeplab_instance']):
            assert sorted(dg) == sorted(int(k) for k in dr), c
            for k, v in dr.items():
                assert abs(dg[int(k)] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (c, k)

@pytest.mark.parametrize('seed', range(MORE or 4))
def test_wrapped_instance_ids_equal_live_reference(seed, ref):
    """More than 255 centres (a lattice of exactly tied peaks): the reference's uint8 ids wrap
    (instance.py:231-236) and its meta dict keeps every centre (:253-266).  The oracle under
    `allow_wrap` must give the same maps, id dicts, meta dicts and orientations, for random
    NMS windows, thresholds, foreground-masked centres and offset scales."""
    from nicr_mt_scene_analysis_b200 import testing
    rng = np.random.default_rng(9100 + seed)
    B, C = 2, int(rng.integers(3, 8))
    H, W = int(rng.integers(66, 90)), int(rng.integers(70, 110))
    ks = int(rng.choice([1, 3, 3, 5]))
    step = int(rng.choice([3, 4])) if ks <= 3 else 4
    apply_fg = bool(rng.integers(0, 2))
    dist_thr = None if rng.integers(0, 2) else int(rng.integers(3, 40))
    data = testing.make_batch(B, C, H, W, 4, seed=500 + seed, quantize='q10', with_orientation=True)
    testing.saturate_heat(data['heat'], step=step, value=float(rng.choice([1.0, 0.5])))
    data['offset'][1] *= float(rng.choice([0.0, 0.01, 0.3]))
    is_thing = tuple(bool(x) for x in rng.integers(0, 2, C))
    if not any(is_thing):
        is_thing = (True,) + is_thing[1:]
    if apply_fg:            # (centres outside the thing mask are dropped: keep enough of them)
        is_thing = (True,) * C
    has_ori = tuple(bool(t and rng.integers(0, 2)) for t in is_thing)
    get = ref['get']
    pan = get('panoptic', semantic_postprocessing=get('semantic')(),
              instance_postprocessing=get(
                  'instance', heatmap_nms_kernel_size=ks, top_k_instances=int(rng.integers(1, 65)),
                  heatmap_apply_foreground_mask=apply_fg, offset_distance_threshold=dist_thr)(),
              semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori)()
    top_k = pan._instance_postprocessing._top_k_instances
    inst_out = (data['heat'], data['offset'], data['orientation'])
    r = pan.postprocess(((data['logits'].clone(), tuple(t.clone() for t in inst_out)), (None, None)),
                        testing.make_batch_dict(B, H, W), is_training=False)
    rmeta = r['panoptic_segmentation_deeplab_instance_meta']
    assert max(len(m) for m in rmeta) > 255, [len(m) for m in rmeta]
    with oracle.allow_wrap():
        got = oracle.panoptic_postprocess(
            data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(),
            data['orientation'].numpy(), is_thing, has_ori, nms_kernel_size=ks, top_k=top_k,
            apply_foreground_mask=apply_fg, offset_distance_threshold=dist_thr, cap=4096)
    assert np.array_equal(got['instance_idx'], r['panoptic_segmentation_deeplab_instance_idx'].numpy())
    assert np.array_equal(got['panoptic'], r['panoptic_segmentation_deeplab'].numpy())
    assert got['ids'] == [{int(k): int(v) for k, v in d.items()}
                          for d in r['panoptic_segmentation_deeplab_ids']]
    for gm, rm in zip(got['meta'], rmeta):
        assert {k: (tuple(v['center_yx']), v['area']) for k, v in gm.items()} == \
            {int(k): (tuple(int(x) for x in v['center_yx']), int(v['area'])) for k, v in rm.items()}
    for dg, dr in zip(got['orientations'], r['orientations_panoptic_segmentation_deeplab_instance']):
        assert sorted(dg) == sorted(int(k) for k in dr)
        for k, v in dr.items():
            assert abs(dg[int(k)] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), k


    # evaluation: the reference's compare_and_accumulate and confusion matrix on the same maps
    L, OFF = 1 << 16, 256 ** 3
    pred = r['panoptic_segmentation_deeplab']
    tgt = torch.roll(pred, int(rng.integers(1, 5)), dims=-1).contiguous()
    for b in range(B):
        try:
            want = ref['pq'](pred[b], tgt[b], C + 1, 0, L, OFF, 0)
        except ZeroDivisionError:           # union == 0 (pq.py:145): the oracle reports code -3
            with pytest.raises(oracle.OracleError) as err:
                oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), C + 1, 0, L, OFF, 0)
            assert err.value.code == -3
            continue
        have = oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), C + 1, 0, L, OFF, 0)
        for w, h in zip(want[:4], have[:4]):
            assert np.array_equal(np.asarray(w, np.float64), h), c            # float64, bit for bit
        assert {(int(g), int(p)) for g, p in want[4]} == have[4], c
    m = ref['miou'](n_classes=C + 1, ignore_first_class=True)
    m.reset()
    m.update(preds=pred // L, target=(tgt // L).to(torch.uint8))
    assert np.array_equal(m.confmat.numpy(), oracle.confmat((pred // L).numpy(), (tgt // L).numpy(), C + 1)), c


def _blocky(rng, B, H, W, n_values, block):
    low = rng.integers(0, n_values, (B, (H + block - 1) // block, (W + block - 1) // block))
    return np.repeat(np.repeat(low, block, 1), block, 2)[:, :H, :W].copy()


@pytest.mark.parametrize('seed', range(MORE or 6))
def test_merges_equal_live_reference(seed, ref):
    """Stand-alone merges on random noisy maps: deeplab_merge_batch (panoptic_merge.py:18-40,
    majority votes with close calls and ties, a foreground mask that disagrees with the
    semantic map, odd id geometries) and naive_merge_semantic_and_instance_np
    (panoptic_merge.py:43-107, uint16 instance ids, instances spanning several classes)."""
    from nicr_mt_scene_analysis.utils.panoptic_merge import (deeplab_merge_batch,
                                                             naive_merge_semantic_and_instance_np)
    rng = np.random.default_rng(8100 + seed)
    B, H, W = int(rng.integers(1, 4)), int(rng.integers(20, 70)), int(rng.integers(20, 90))
    n_sem = int(rng.integers(3, 12))
    sem = _blocky(rng, B, H, W, n_sem, int(rng.integers(4, 12)))
    noise = rng.integers(0, n_sem, (B, H, W))
    sem = np.where(rng.random((B, H, W)) < 0.3, noise, sem)                    # 0 = void
    ins = _blocky(rng, B, H, W, int(rng.integers(2, 12)), int(rng.integers(5, 16))).astype(np.uint8)
    fg = _blocky(rng, B, H, W, 4, int(rng.integers(3, 9))) > 0
    thing_ids = np.flatnonzero(rng.integers(0, 2, n_sem))
    thing_ids = thing_ids[thing_ids > 0]
    L = int(rng.choice([1 << 16, 1000, 256]))
    void = int(rng.choice([0, 0, int(rng.integers(0, n_sem))]))
    want, want_ids = deeplab_merge_batch(torch.from_numpy(sem), torch.from_numpy(ins),
                                         torch.from_numpy(fg), L, thing_ids, void)
    have, have_ids = oracle.deeplab_merge_batch(sem, ins, fg, L, thing_ids.tolist(), void)
    assert np.array_equal(have, want.numpy())
    assert have_ids == [{int(k): int(v) for k, v in d.items()} for d in want_ids]

    ins16 = _blocky(rng, B, H, W, 12, int(rng.integers(6, 16))).astype(np.uint16)
    ins16[ins16 == 7] = 40000
    ins16[ins16 == 11] = 65535
    sem8 = sem.astype(np.uint8)
    have, have_ids = oracle.naive_merge_batch(sem8, ins16.astype(np.int32), 1 << 16,
                                              thing_ids.tolist(), 0)
    for b in range(B):
        want, d = naive_merge_semantic_and_instance_np(sem8[b], ins16[b], 1 << 16, thing_ids, 0)
        assert np.array_equal(have[b], want.astype(np.int64))
        assert have_ids[b] == {int(k): int(v) for k, v in d.items()}


@pytest.mark.parametrize('seed', range(MORE or 4))
def test_instance_targets_equal_live_reference(seed, ref):
    """InstanceTargetGenerator (data/preprocessing/instance.py:97-286) on random ground truth:
    float32 centre heat-maps bit for bit, offsets (normalised and in pixels), masks."""
    from nicr_mt_scene_analysis.data.preprocessing.instance import InstanceTargetGenerator
    rng = np.random.default_rng(8200 + seed)
    H, W = int(rng.integers(30, 80)), int(rng.integers(30, 110))
    sigma = int(rng.choice([3, 5, 8]))
    n_cls = int(rng.integers(3, 8))
    is_thing = (False,) + tuple(bool(x) for x in rng.integers(0, 2, n_cls - 1))
    ins = _blocky(rng, 1, H, W, 9, int(rng.integers(7, 16)))[0].astype(np.uint16)
    ins[ins == 5] = 51234
    sem = np.zeros((H, W), np.uint8)
    for i in np.unique(ins):
        sem[ins == i] = int(rng.integers(1, n_cls))
    noise = rng.random((H, W)) < 0.2
    sem[noise] = rng.integers(0, n_cls, int(noise.sum())).astype(np.uint8)
    # the reference asserts that no stuff pixel carries an instance id
    ins[~np.asarray(is_thing)[sem]] = 0
    for norm in (True, False):
        gen = InstanceTargetGenerator(sigma=sigma, semantic_classes_is_thing=is_thing,
                                      normalized_offset=norm)
        want = gen({'semantic': sem.copy(), 'instance': ins.copy()})
        have = oracle.instance_targets(sem[None], ins[None].astype(np.int32), sigma, list(is_thing), norm)
        assert np.array_equal(have['instance_center'][0], want['instance_center'])
        assert np.array_equal(have['instance_offset'][0], want['instance_offset'].transpose(2, 0, 1))
        assert np.array_equal(have['instance_foreground'][0], want['instance_foreground'])
        assert np.array_equal(have['instance_center_mask'][0], want['instance_center_mask'])


@pytest.mark.parametrize('seed', range(MORE or 10))
def test_pq_frames_equal_live_reference(seed, ref):
    """compare_and_accumulate (metric/pq.py:60-179) on random blocky / noisy panoptic maps with
    odd id geometries: any ignored label, offsets that are not powers of two, predictions inside
    void and ignored ground truth, thin and single-pixel segments.  float64 bit for bit."""
    rng = np.random.default_rng(9300 + seed)
    H, W = int(rng.integers(8, 60)), int(rng.integers(8, 80))
    NC = int(rng.integers(2, 12))
    L = int(rng.choice([1 << 16, 1000, 37]))
    n_inst = int(rng.integers(1, min(L, 9)))
    ignored = int(rng.integers(0, NC))
    offset = int(rng.choice([256 ** 3, 3 * 10 ** 7, NC * L + 1]))
    assert offset > NC * L

    def random_map():
        cat = _blocky(rng, 1, H, W, NC, int(rng.integers(2, 12)))[0]
        ins = _blocky(rng, 1, H, W, n_inst, int(rng.integers(2, 9)))[0]
        noise = rng.random((H, W)) < rng.choice([0.0, 0.05, 0.3])
        cat = np.where(noise, rng.integers(0, NC, (H, W)), cat)
        return (cat * L + ins).astype(np.int64)

    tgt = random_map()
    pred = np.where(rng.random((H, W)) < 0.7, tgt, random_map())
    void_segment_id = ignored * L
    try:
        want = ref['pq'](torch.from_numpy(pred), torch.from_numpy(tgt), NC, ignored, L, offset,
                         void_segment_id)
    except ZeroDivisionError:               # union == 0 (pq.py:145): the oracle reports code -3
        with pytest.raises(oracle.OracleError) as err:
            oracle.pq_compare_and_accumulate(pred, tgt, NC, ignored, L, offset, void_segment_id)
        assert err.value.code == -3
        return
    have = oracle.pq_compare_and_accumulate(pred, tgt, NC, ignored, L, offset, void_segment_id)
    for w, h in zip(want[:4], have[:4]):
        assert np.array_equal(np.asarray(w, np.float64), h), (NC, L, ignored, offset)
    assert {(int(g), int(p)) for g, p in want[4]} == have[4]


@pytest.mark.parametrize('seed', range(MORE or 6))
def test_instance_stage_functions_equal_live_reference(seed, ref):
    """InstancePostprocessing._get_instance_segmentation / _get_instance_orientation called
    directly with an ARBITRARY foreground mask (the ground-truth foreground path,
    instance.py:371-449, calls them that way): instance map and meta identical, angles 1e-5."""
    from nicr_mt_scene_analysis_b200 import testing
    rng = np.random.default_rng(9400 + seed)
    B, H, W = int(rng.integers(1, 3)), int(rng.integers(24, 70)), int(rng.integers(24, 90))
    K = int(rng.integers(1, 7))
    data = testing.make_batch(B, 4, H, W, K, seed=500 + seed, quantize=str(rng.choice(['q10', 'tie'])),
                              with_orientation=True)
    normalized = bool(rng.integers(0, 2))
    thr, ks, top_k = float(rng.choice([0.1, 0.3])), int(rng.choice([3, 5])), int(rng.integers(1, 9))
    apply_fg = bool(rng.integers(0, 2))
    dist_thr = None if rng.integers(0, 2) else int(rng.integers(3, 25))
    fg = torch.from_numpy(_blocky(rng, B, H, W, 3, int(rng.integers(3, 12))) > 0)
    offset = data['offset'].clone()
    if not normalized:
        offset[:, 0] *= H
        offset[:, 1] *= W
    ins = ref['get']('instance', heatmap_threshold=thr, heatmap_nms_kernel_size=ks,
                     top_k_instances=top_k, heatmap_apply_foreground_mask=apply_fg,
                     normalized_offset=normalized, offset_distance_threshold=dist_thr)()
    scaled = offset.clone()                  # the caller de-normalises (panoptic.py:105-111)
    if normalized:
        scaled[:, 0] *= H
        scaled[:, 1] *= W
    want_seg, want_meta = ins._get_instance_segmentation(data['heat'], scaled, fg)
    have_seg, have_meta = oracle.instance_segmentation(
        data['heat'].numpy(), offset.numpy(), fg.numpy(), thr, ks, top_k, apply_fg, normalized, dist_thr)
    assert np.array_equal(have_seg, want_seg.numpy())
    for gm, rm in zip(have_meta, want_meta):
        assert {k: (tuple(v['center_yx']), v['area']) for k, v in gm.items()} == \
            {int(k): (tuple(int(x) for x in v['center_yx']), int(v['area'])) for k, v in rm.items()}
    ori_mask = torch.from_numpy(_blocky(rng, B, H, W, 2, int(rng.integers(3, 12))) > 0)
    for mask in (ori_mask, None):
        want = ins._get_instance_orientation(data['orientation'], want_seg, mask)
        have = oracle.instance_orientation(data['orientation'].numpy(), want_seg.numpy(),
                                           None if mask is None else mask.numpy())
        assert [sorted(int(k) for k in d) for d in want] == [sorted(d) for d in have]
        for dw, dh in zip(want, have):
            for k, v in dw.items():
                assert abs(dh[int(k)] - v) <= 1e-5 * max(1.0, abs(v)), (k, v, dh[int(k)])


def test_product_compute_equals_live_reference(ref):
    """The PRODUCT's host arithmetic (not the oracle): `PanopticQuality.compute` /
    `result_per_category` and `MeanIntersectionOverUnion.compute` of this package on hand-set
    states against the reference's own metric objects holding the same states (pq.py:254-361,
    miou.py:58-94): every key, every value."""
    from nicr_mt_scene_analysis.metric import PanopticQuality as RefPQ
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticQuality
    rng = np.random.default_rng(77)
    for trial in range(12):
        n = int(rng.integers(2, 14))
        ignored = int(rng.integers(0, n))
        is_thing = [bool(x) for x in rng.integers(0, 2, n)]
        tp = rng.integers(0, 6, n).astype(np.float64) * (rng.random(n) < 0.7)
        fn = rng.integers(0, 5, n).astype(np.float64) * (rng.random(n) < 0.6)
        fp = rng.integers(0, 5, n).astype(np.float64) * (rng.random(n) < 0.6)
        iou = tp * rng.uniform(0.5, 1.0, n)
        theirs = RefPQ(n, ignored, 1 << 16, 256 ** 3, is_thing, num_workers=1)
        ours = PanopticQuality(n, ignored, 1 << 16, 256 ** 3, is_thing, device='cpu')
        for m in (theirs, ours):
            m.iou_per_class = torch.from_numpy(iou.copy())
            m.tp_per_class = torch.from_numpy(tp.copy())
            m.fn_per_class = torch.from_numpy(fn.copy())
            m.fp_per_class = torch.from_numpy(fp.copy())
        want, got = theirs.compute(suffix='_s'), ours.compute(suffix='_s')
        assert set(want) == set(got), trial
        for k, v in want.items():
            np.testing.assert_array_equal(np.asarray(torch.as_tensor(got[k]).double()),
                                          np.asarray(torch.as_tensor(v).double()), err_msg=f'{trial} {k}')
        want, got = theirs.result_per_category(), ours.result_per_category()
        assert set(want) == set(got)
        for k, v in want.items():
            np.testing.assert_array_equal(got[k].numpy(), v.numpy(), err_msg=f'{trial} {k}')
        del theirs                                  # terminates its worker pool

        cm = rng.integers(0, 50, (n, n)) * (rng.random((n, n)) < 0.6)
        cm[int(rng.integers(0, n))] = 0             # a class without ground truth
        for flag in (False, True):
            theirs = ref['miou'](n_classes=n, ignore_first_class=flag)
            ours_m = MeanIntersectionOverUnion(n, ignore_first_class=flag, device='cpu')
            theirs.confmat = torch.from_numpy(cm.astype(np.int64))
            ours_m.confmat = torch.from_numpy(cm.astype(np.int64))
            (w_miou, w_ious), (g_miou, g_ious) = theirs.compute(return_ious=True), ours_m.compute(return_ious=True)
            np.testing.assert_array_equal(g_ious.numpy(), w_ious.numpy())
            assert (torch.isnan(w_miou) and torch.isnan(g_miou)) or float(g_miou) == float(w_miou)
            assert float(ours_m.compute()) == float(theirs.compute()) or torch.isnan(theirs.compute())


def test_product_angular_errors_equal_live_reference(ref):
    """MeanAbsoluteAngularError.update and PanopticQualityWithOrientationMAE.update_mae of this
    package (one vector operation per batch) against the reference's per-pair loops
    (mae.py:46-58, 129-162) on random dicts: same element count, float64 sums equal (bit for
    bit where the visiting order is the same, 1e-12 for the set-ordered matches)."""
    from nicr_mt_scene_analysis.metric.mae import (MeanAbsoluteAngularError as RefMAE,
                                                   PanopticQualityWithOrientationMAE as RefPQMAE)
    from nicr_mt_scene_analysis_b200.metric import (MeanAbsoluteAngularError,
                                                    PanopticQualityWithOrientationMAE)
    rng = np.random.default_rng(91)
    preds = [{int(i): float(rng.uniform(-7, 7)) for i in rng.choice(50, int(rng.integers(0, 20)), replace=False)}
             for _ in range(6)]
    targets = [dict({k: float(rng.uniform(-7, 7)) for k in d}, extra=1.0) for d in preds]
    theirs, ours = RefMAE(), MeanAbsoluteAngularError(device='cpu')
    theirs.update(preds, targets)
    ours.update(preds, targets)
    assert int(ours.n_elements) == int(theirs.n_elements) == sum(len(d) for d in preds)
    assert float(ours.sum_angular_error) == float(theirs.sum_angular_error)
    for a, b in zip(ours.compute(), theirs.compute()):
        assert float(a) == float(b)

    L = 1 << 16
    theirs = RefPQMAE(4, 0, L, 256 ** 3, [False, True, True, False], num_workers=1)
    ours = PanopticQualityWithOrientationMAE(4, 0, L, 256 ** 3, [False, True, True, False], device='cpu')
    for _ in range(5):
        gt_ids = [int(c * L + i) for c in (1, 2) for i in range(1, 6)]
        pred_ids = [int(c * L + i) for c in (1, 2) for i in range(1, 7)]
        matching = {(g, p) for g, p in zip(gt_ids, rng.permutation(pred_ids)[:len(gt_ids)].tolist())}
        matching |= {(0, 0), (3 * L, 3 * L)}                      # stuff / void pairs
        tgt_id_dict = {g: g % L + 100 for g in gt_ids if rng.random() < 0.8}
        pred_id_dict = {p: p % L + 200 for p in pred_ids if rng.random() < 0.8}
        ori_t = {v: float(rng.uniform(-4, 4)) for v in tgt_id_dict.values() if rng.random() < 0.7}
        ori_p = {v: float(rng.uniform(-4, 4)) for v in pred_id_dict.values() if rng.random() < 0.7}
        theirs.update_mae(ori_p, pred_id_dict, ori_t, tgt_id_dict, matching)
        ours.update_mae(ori_p, pred_id_dict, ori_t, tgt_id_dict, matching)
    assert int(ours.n_elements) == int(theirs.n_elements) > 0
    np.testing.assert_allclose(float(ours.sum_angular_error), float(theirs.sum_angular_error), rtol=1e-12)
    del theirs


@pytest.mark.parametrize('seed', range(MORE or 6))
def test_nonfinite_logits_equal_live_reference(seed, ref):
    """NaN / +-Inf logits: the reference's `softmax -> max` (semantic.py:52-53) answers class 0
    wherever a NaN or +Inf poisons the soft-max (or nothing but -Inf is there) and ignores -Inf
    next to finite logits; the whole panoptic result follows from that class map."""
    from nicr_mt_scene_analysis_b200 import testing
    rng = np.random.default_rng(9100 + seed)
    B, C, H, W, K = int(rng.integers(1, 3)), int(rng.integers(1, 12)), int(rng.integers(16, 56)), \
        int(rng.integers(16, 72)), int(rng.integers(1, 5))
    data = testing.make_batch(B, C, H, W, K, seed=500 + seed, quantize='q10', with_orientation=False)
    testing.poison_logits(data['logits'], float(rng.choice([0.02, 0.2, 0.6])), seed)
    is_thing = tuple(bool(x) for x in rng.integers(0, 2, C))
    get = ref['get']
    pan = get('panoptic', semantic_postprocessing=get('semantic')(),
              instance_postprocessing=get('instance')(), semantic_classes_is_thing=is_thing,
              semantic_class_has_orientation=(False,) * C)()
    r = pan.postprocess(((data['logits'].clone(), (data['heat'].clone(), data['offset'].clone())),
                         (None, None)), testing.make_batch_dict(B, H, W), is_training=False)
    want = r['semantic_segmentation_idx'].numpy()
    assert np.array_equal(oracle.semantic_argmax(data['logits'].numpy()), want)
    got = oracle.panoptic_postprocess(data['logits'].numpy(), data['heat'].numpy(),
                                      data['offset'].numpy(), None, is_thing, (False,) * C)
    assert np.array_equal(got['panoptic'], r['panoptic_segmentation_deeplab'].numpy())
    assert np.array_equal(got['instance_idx'], r['panoptic_segmentation_deeplab_instance_idx'].numpy())
    np.testing.assert_allclose(oracle.semantic_score(data['logits'].numpy()),
                               r['semantic_segmentation_score'].numpy(), rtol=1e-5)


UNQUANTISED_FRAMES = 14     # x 480 x 640 = 4.3 M pixels of the bench distribution


def test_unquantised_bench_inputs_have_no_softmax_flips(ref):
    """bench.py feeds UNQUANTISED logits (testing.make_frame(quantize=None), seeds 1000 + i of
    rank 0).  The product and the oracle take the arg-max of the logits, the reference the arg-max
    of softmax(logits); they can only differ where two logits are closer than ~2^-23 relative.
    On 4.3 M pixels of exactly the frames the bench uses there is no such pixel (0 flips);
    tests/test_gpu_postprocessing.py::test_unquantised_bench_frames checks the CUDA path against
    the oracle on the same frames."""
    import torch.nn.functional as F
    from nicr_mt_scene_analysis_b200 import testing
    flips = 0
    for i in range(UNQUANTISED_FRAMES):
        f = testing.make_frame(40, 480, 640, 12, seed=1000 + i, with_orientation=False, quantize=None)
        want = torch.max(F.softmax(f['logits'][None], dim=1), dim=1)[1][0].numpy()
        have = oracle.semantic_argmax(f['logits'][None].numpy())[0]
        flips += int((want != have).sum())
    assert flips == 0

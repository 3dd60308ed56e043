"""The C oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

import oracle
from conftest import int_keys, jload, load_golden

POST_CASES = ['q10', 'tie', 'odd', 'pixel_offsets', 'scores', 'nonfinite']


def _meta_equal(ref_meta, got_meta, orient=None):
    assert len(ref_meta) == len(got_meta)
    for rm, gm in zip(ref_meta, got_meta):
        rm = {int(k): v for k, v in rm.items()}
        assert sorted(rm) == sorted(gm)
        for i in rm:
            assert tuple(rm[i]['center_yx']) == tuple(gm[i]['center_yx'])
            assert rm[i]['area'] == gm[i]['area']
            assert rm[i]['score'] == pytest.approx(gm[i]['score'], rel=1e-6)


@pytest.mark.parametrize('name', POST_CASES)
def test_postprocess_pipeline(name):
    z = load_golden('post_' + name)
    cfg = jload(z['cfg'])
    r = oracle.panoptic_postprocess(
        z['logits'], z['heat'], z['offset'], z.get('orientation'), z['is_thing'],
        z['has_orientation'], threshold=cfg['thr'], nms_kernel_size=cfg['ks'],
        top_k=cfg['top_k'], apply_foreground_mask=cfg['apply_fg'],
        normalized_offset=cfg['normalized'], offset_distance_threshold=cfg['dist_thr'])
    assert np.array_equal(r['semantic_idx'], z['semantic_segmentation_idx'])
    assert np.array_equal(r['instance_idx'], z['panoptic_segmentation_deeplab_instance_idx'])
    assert np.array_equal(r['panoptic'], z['panoptic_segmentation_deeplab'])
    assert r['ids'] == int_keys(jload(z['ids']))
    _meta_equal(jload(z['meta']), r['meta'])
    if 'orientations' in z:
        ref = int_keys(jload(z['orientations']))
        assert [sorted(d) for d in ref] == [sorted(d) for d in r['orientations']]
        for dr, dg in zip(ref, r['orientations']):
            for k in dr:
                assert dg[k] == pytest.approx(dr[k], rel=1e-5, abs=1e-6)


def test_postprocess_pipeline_with_wrapped_instance_ids():
    """`post_wrap`: two frames with 432 exactly tied heat-map peaks.  The reference keeps all of
    them and its uint8 ids wrap (instance.py:236): centre 256 is "no instance", centre 257 joins
    instance 1, the meta dict has 432 entries (zero areas beyond 255).  The oracle refuses such
    frames by default and reproduces the reference under `allow_wrap`."""
    z = load_golden('post_wrap')
    cfg = jload(z['cfg'])
    args = (z['logits'], z['heat'], z['offset'], z.get('orientation'), z['is_thing'],
            z['has_orientation'])
    kw = dict(threshold=cfg['thr'], nms_kernel_size=cfg['ks'], top_k=cfg['top_k'],
              apply_foreground_mask=cfg['apply_fg'], normalized_offset=cfg['normalized'],
              offset_distance_threshold=cfg['dist_thr'])
    with pytest.raises(oracle.OracleError) as err:
        oracle.panoptic_postprocess(*args, **kw)
    assert err.value.code == -2
    with oracle.allow_wrap():
        r = oracle.panoptic_postprocess(*args, cap=1024, **kw)
    assert [len(m) for m in r['meta']] == [5, 432, 432]
    assert np.array_equal(r['instance_idx'], z['panoptic_segmentation_deeplab_instance_idx'])
    assert np.array_equal(r['panoptic'], z['panoptic_segmentation_deeplab'])
    assert r['ids'] == int_keys(jload(z['ids']))
    _meta_equal(jload(z['meta']), r['meta'])
    ref = int_keys(jload(z['orientations']))
    assert [sorted(d) for d in ref] == [sorted(d) for d in r['orientations']]
    for dr, dg in zip(ref, r['orientations']):
        for k in dr:
            assert dg[k] == pytest.approx(dr[k], rel=1e-5, abs=1e-6)


@pytest.mark.parametrize('name', POST_CASES)
def test_stage_functions(name):
    """stage-wise: arg-max, centres, grouping, merge each against the reference."""
    z = load_golden('post_' + name)
    cfg = jload(z['cfg'])
    sem = oracle.semantic_argmax(z['logits'])
    assert np.array_equal(sem, z['semantic_segmentation_idx'])
    fg = z['is_thing'][sem]
    assert np.array_equal(fg, z['panoptic_foreground_mask'])
    mask, centers = oracle.instance_centers(z['heat'], cfg['thr'], cfg['ks'], cfg['top_k'],
                                            fg, cfg['apply_fg'])
    assert np.array_equal(mask, z['center_mask'])
    assert [c.tolist() for c in centers] == jload(z['centers'])
    inst, meta = oracle.instance_segmentation(
        z['heat'], z['offset'], fg, cfg['thr'], cfg['ks'], cfg['top_k'], cfg['apply_fg'],
        cfg['normalized'], cfg['dist_thr'])
    assert np.array_equal(inst, z['panoptic_segmentation_deeplab_instance_idx'])
    _meta_equal(jload(z['meta']), meta)
    thing_ids = (np.nonzero(z['is_thing'])[0] + 1).tolist()
    pan, ids = oracle.deeplab_merge_batch(sem.astype(np.int32) + 1, inst, fg, 1 << 16,
                                          thing_ids, 0)
    assert np.array_equal(pan, z['panoptic_segmentation_deeplab'])
    assert ids == int_keys(jload(z['ids']))
    score = oracle.semantic_score(z['logits'])
    np.testing.assert_allclose(score, z['semantic_segmentation_score'], rtol=1e-5)


def test_centers_tie_cases():
    cases = jload(load_golden('centers')['cases'])
    assert len(cases) == 12
    for c in cases:
        heat = np.array(c['heat'], np.float32)
        fg = np.array(c['fg'], np.uint8)
        mask, centers = oracle.instance_centers(heat, c['thr'], c['ks'], c['k'], fg,
                                                c['apply_fg'])
        assert np.array_equal(mask, np.array(c['mask'], bool)), c['ks']
        assert [x.tolist() for x in centers] == c['centers']


def test_merge_standalone():
    z = load_golden('merge')
    pan, ids = oracle.deeplab_merge_batch(z['sem'], z['ins'], z['fg'], int(z['L']),
                                          z['thing_ids'].tolist(), 0)
    assert np.array_equal(pan, z['pan'])
    assert ids == int_keys(jload(z['ids']))
    pan, ids = oracle.deeplab_merge_batch(z['sem'], z['ins'], z['fg'], 1000,
                                          z['thing_ids'].tolist(), 3)
    assert np.array_equal(pan, z['pan_L1000_void3'])
    assert ids == int_keys(jload(z['ids_L1000_void3']))


def test_pq_frames_bit_exact():
    z = load_golden('pq')
    matches = jload(z['matches'])
    state = np.zeros((4, int(z['num_categories'])), np.float64)
    for b in range(z['pred'].shape[0]):
        iou, tp, fn, fp, m = oracle.pq_compare_and_accumulate(
            z['pred'][b], z['target'][b], int(z['num_categories']), 0, int(z['L']),
            int(z['offset']), 0)
        # float64 IoU sums are compared BIT-exactly (same visiting order as pq.py:119)
        assert np.array_equal(iou, z['iou'][b])
        assert np.array_equal(tp, z['tp'][b])
        assert np.array_equal(fn, z['fn'][b])
        assert np.array_equal(fp, z['fp'][b])
        assert sorted([list(x) for x in m]) == matches[b]
        for s, v in zip(state, (iou, tp, fn, fp)):
            s += v
    assert np.array_equal(state, z['state'])


def _pq_single(pred, tgt, **kw):
    return oracle.pq_compare_and_accumulate(np.array(pred), np.array(tgt), **kw)


def test_pq_known_answers():
    """the reference's own known-answer cases, tests/test_metrics.py:76-446"""
    inst = np.array([[1, 1, 1, 1, 1, 1], [1, 2, 2, 2, 2, 1], [1, 2, 2, 2, 2, 1],
                     [1, 2, 2, 2, 2, 1], [1, 2, 2, 1, 1, 1], [1, 2, 1, 1, 1, 1]])
    kw = dict(num_categories=1, ignored_label=2, max_instances_per_category=16, offset=16,
              void_segment_id=32)
    iou, tp, fn, fp, _ = _pq_single(inst, inst, **kw)            # perfect match  :76-113
    assert (iou, tp, fn, fp) == ([2.0], [2], [0], [0])
    cat = np.array([[0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 1, 0], [0, 1, 1, 1, 1, 0],
                    [0, 1, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]])
    iou, tp, fn, fp, _ = _pq_single(1 - cat, cat, num_categories=2, ignored_label=2,  # :116-161
                                    max_instances_per_category=1, offset=16, void_segment_id=2)
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == ([0, 0], [0, 0], [1, 1], [1, 1])
    gt = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 2, 2, 2, 1],
                   [1, 2, 2, 2, 2, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]])
    good = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 2, 2, 2, 2, 1],
                     [1, 2, 2, 2, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]])
    bad = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 1, 1, 2, 2, 1],
                    [1, 1, 1, 2, 2, 1], [1, 1, 1, 2, 2, 1], [1, 1, 1, 1, 1, 1]])
    iou, tp, fn, fp, _ = _pq_single(good, gt, **kw)              # matches by iou  :164-257
    assert iou[0] == pytest.approx(28 / 30 + 6 / 8) and (tp, fn, fp) == ([2], [0], [0])
    iou, tp, fn, fp, _ = _pq_single(bad, gt, **kw)
    assert iou[0] == pytest.approx(27 / 32) and (tp, fn, fp) == ([1], [1], [1])
    cat = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1], [1, 2, 2, 1, 2, 2],
                    [1, 2, 2, 1, 2, 2], [1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]])
    pinst = np.zeros((6, 6), int)
    pinst[2:4, 4:6] = 1
    kw3 = dict(num_categories=3, ignored_label=0, max_instances_per_category=10, offset=100,
               void_segment_id=0)
    iou, tp, fn, fp, _ = _pq_single(cat * 10 + pinst, cat * 10, **kw3)   # wrong instances :260-316
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == \
        ([0, 1, 0], [0, 1, 0], [0, 0, 1], [0, 0, 2])
    ginst = np.zeros((6, 6), int)
    ginst[2:4, 1:3] = 1
    iou, tp, fn, fp, m = _pq_single(cat * 10 + pinst, cat * 10 + ginst, **kw3)  # :319-381
    assert (iou.tolist(), tp.tolist(), fn.tolist(), fp.tolist()) == \
        ([0, 1, 2], [0, 1, 2], [0, 0, 0], [0, 0, 0])
    res = oracle.pq_results(iou, tp, fn, fp, [True] * 3, 0)
    assert res['all_pq'] == 1.0 and res['all_num_categories'] == 2


def test_pq_zero_division_is_reported():
    # gt void (cat 0 == ignored) fully covered by a cat-0 prediction: union = 0 (pq.py:143-145)
    with pytest.raises(oracle.OracleError) as e:
        _pq_single(np.full((4, 4), 3), np.zeros((4, 4), int), num_categories=2, ignored_label=0,
                   max_instances_per_category=10, offset=100, void_segment_id=0)
    assert e.value.code == -3


@pytest.mark.parametrize('n', [6, 41, 200])
def test_miou(n):
    z = load_golden('miou')
    cm = np.zeros((n, n), np.int64)
    for p, t in zip(z[f'pred_{n}'], z[f'target_{n}']):
        cm += oracle.confmat(p, t, n)
    assert np.array_equal(cm, z[f'confmat_{n}'])
    for flag in (0, 1):
        miou, ious = oracle.miou_from_confmat(cm, bool(flag))
        # torch.mean's f32 summation order is not restated: last-bit tolerance on the mean
        assert miou == pytest.approx(float(z[f'miou{flag}_{n}']), rel=1e-6)
        assert np.array_equal(ious, z[f'ious{flag}_{n}'], equal_nan=True)


def test_orientation_standalone():
    z = load_golden('orientation')
    for key, mask in (('with_mask', z['mask']), ('without_mask', None)):
        ref = int_keys(jload(z[key]))
        got = oracle.instance_orientation(z['ori'], z['seg'], mask)
        assert [sorted(d) for d in ref] == [sorted(d) for d in got]
        for dr, dg in zip(ref, got):
            for k in dr:
                assert math.isclose(dg[k], dr[k], rel_tol=1e-5, abs_tol=1e-6)


def test_naive_merge_ground_truth_targets():
    z = load_golden('naive_merge')
    pan, ids = oracle.naive_merge_batch(z['sem'], z['ins'], 1 << 16, z['thing_ids'].tolist(), 0)
    assert np.array_equal(pan, z['pan'])
    assert ids == int_keys(jload(z['ids']))
    assert [list(d) for d in ids] == [[int(k) for k in d] for d in jload(z['ids'])]   # creation order


def test_instance_targets():
    z = load_golden('instance_targets')
    r = oracle.instance_targets(z['sem'], z['ins'], 5, z['is_thing'].tolist(), True)
    assert np.array_equal(r['instance_center'], z['instance_center'])            # bit-exact f32
    assert np.array_equal(r['instance_offset'], z['instance_offset'].transpose(0, 3, 1, 2))
    assert np.array_equal(r['instance_foreground'], z['instance_foreground'])
    assert np.array_equal(r['instance_center_mask'], z['instance_center_mask'])
    r = oracle.instance_targets(z['sem'], z['ins'], 5, z['is_thing'].tolist(), False)
    assert np.array_equal(r['instance_offset'], z['instance_offset_px'].transpose(0, 3, 1, 2))
    bad = z['ins'].copy()
    bad[0, 0, 0] = 999 if z['sem'][0, 0, 0] in (0, 2, 5) else bad[0, 0, 0]
    stuff_px = np.argwhere(~z['is_thing'][z['sem'][0]])[0]
    bad[0, stuff_px[0], stuff_px[1]] = 999          # a lone stuff pixel with an instance id
    with pytest.raises(AssertionError):
        oracle.instance_targets(z['sem'], bad, 5, z['is_thing'].tolist(), True)


def test_task_helper_epoch_results():
    """The oracle reproduces what the reference's PanopticTaskHelper / InstanceTaskHelper log
    at the end of a validation epoch (tests/golden/task_helpers.npz): PQ states accumulated
    over two steps, the confusion matrix of `pred // L`, and for the instance helper the merge
    of GT semantic + predicted instances that precedes the PQ."""
    z = load_golden('task_helpers')
    NC, L, OFF = int(z['num_categories']), 1 << 16, 256 ** 3
    is_thing = z['is_thing']
    thing_ids = [int(i) for i in np.nonzero(is_thing)[0]]
    st_pan, st_ins = np.zeros((4, NC)), np.zeros((4, NC))
    cm = np.zeros((NC, NC), np.int64)
    for i in range(int(z['n_steps'])):
        pan_t, pan_p = z[f'step{i}/pan_t'], z[f'step{i}/pan_p']
        merged, _ = oracle.deeplab_merge_batch(z[f'step{i}/sem_t'], z[f'step{i}/inst_fg'],
                                               z[f'step{i}/inst_gt'] != 0, L, thing_ids, 0)
        for b in range(pan_t.shape[0]):
            for st, pred in ((st_pan, pan_p[b]), (st_ins, merged[b])):
                out = oracle.pq_compare_and_accumulate(pred, pan_t[b], NC, 0, L, OFF, 0)
                for s, v in zip(st, out[:4]):
                    s += v
        cm += oracle.confmat(pan_p // L, z[f'step{i}/sem_t'], NC)
    assert np.array_equal(cm, z['panoptic/artifacts/panoptic_deeplab_semantic_cm'])
    miou, ious = oracle.miou_from_confmat(cm, True)
    np.testing.assert_allclose(miou, z['panoptic/logs/panoptic_deeplab_semantic_miou'], rtol=1e-6)
    for prefix, st in (('panoptic', st_pan), ('instance', st_ins)):
        res = oracle.pq_results(*st, is_thing, 0, suffix='_deeplab')
        for k, v in res.items():
            kind = 'artifacts' if np.ndim(v) else 'logs'
            np.testing.assert_allclose(np.float64(v), z[f'{prefix}/{kind}/{prefix}_{k}'],
                                       rtol=1e-14, err_msg=k)


def test_semantic_task_helper_epoch_results():
    """The oracle's confusion matrix over the non-void pixels (target - 1) reproduces what the
    reference's SemanticTaskHelper logs (tests/golden/semantic_helper.npz)."""
    z = load_golden('semantic_helper')
    C = int(z['n_classes'])
    cm = np.zeros((C, C), np.int64)
    for i in range(int(z['n_steps'])):
        t, p = z[f'step{i}/target'], z[f'step{i}/preds']
        keep = t != 0
        cm += oracle.confmat(p[keep], t[keep].astype(np.int64) - 1, C)
    assert np.array_equal(cm, z['artifacts/semantic_cm'])
    miou, ious = oracle.miou_from_confmat(cm, False)
    np.testing.assert_allclose(miou, z['logs/semantic_miou'], rtol=1e-6)
    np.testing.assert_allclose(ious, z['artifacts/semantic_ious_per_class'], rtol=1e-6,
                               equal_nan=True)

"""CUDA path vs (a) the golden vectors produced by the unmodified reference and (b) the C
oracle on seeded synthetic inputs.  Everything integer is compared bit-exactly."""
import math

import numpy as np
import pytest
import torch

import oracle
from conftest import int_keys, jload, load_golden

pytestmark = pytest.mark.gpu

POST_CASES = ['q10', 'tie', 'odd', 'pixel_offsets', 'scores', 'nonfinite']


def _build(cfg, is_thing, has_ori, **extra):
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    sem = get_postprocessing_class('semantic')()
    ins = get_postprocessing_class(
        'instance', heatmap_threshold=cfg['thr'], heatmap_nms_kernel_size=cfg['ks'],
        top_k_instances=cfg['top_k'], heatmap_apply_foreground_mask=cfg['apply_fg'],
        normalized_offset=cfg['normalized'], offset_distance_threshold=cfg['dist_thr'])()
    pan = get_postprocessing_class(
        'panoptic', semantic_postprocessing=sem, instance_postprocessing=ins,
        semantic_classes_is_thing=tuple(bool(x) for x in is_thing),
        semantic_class_has_orientation=tuple(bool(x) for x in has_ori),
        normalized_offset=cfg['normalized'], compute_scores=cfg.get('compute_scores', False),
        **extra)()
    return sem, ins, pan


def _run(pan, logits, heat, offset, orientation, dev):
    from nicr_mt_scene_analysis_b200 import testing
    B, _, H, W = logits.shape
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    inst = (t(heat), t(offset)) + ((t(orientation),) if orientation is not None else ())
    return pan.postprocess(((t(logits), inst), (None, None)), testing.make_batch_dict(B, H, W),
                           is_training=False)


def _check_meta(ref_meta, got_meta):
    assert len(ref_meta) == len(got_meta)
    for rm, gm in zip(ref_meta, got_meta):
        rm = {int(k): v for k, v in rm.items()}
        assert sorted(rm) == sorted(gm)
        for i in rm:
            assert tuple(rm[i]['center_yx']) == tuple(gm[i]['center_yx'])
            assert rm[i]['area'] == gm[i]['area']
            assert rm[i]['score'] == pytest.approx(gm[i]['score'], rel=1e-6)
            for k in ('semantic_score', 'panoptic_score'):      # compute_scores=True
                if k in rm[i]:
                    assert gm[i][k] == pytest.approx(rm[i][k], rel=1e-5), (i, k)
            for k in ('semantic_idx', 'panoptic_id'):
                if k in rm[i]:
                    assert gm[i][k] == rm[i][k], (i, k)
            if 'orientation' in rm[i]:
                if math.isnan(rm[i]['orientation']):
                    assert math.isnan(gm[i]['orientation'])
                else:
                    assert math.isclose(gm[i]['orientation'], rm[i]['orientation'],
                                        rel_tol=1e-5, abs_tol=1e-6)


def _check_orient(ref, got):
    assert [sorted(d) for d in ref] == [sorted(d) for d in got]
    for dr, dg in zip(ref, got):
        for k in dr:
            assert math.isclose(dg[k], dr[k], rel_tol=1e-5, abs_tol=1e-6), (k, dg[k], dr[k])


@pytest.mark.parametrize('name', POST_CASES)
@pytest.mark.parametrize('async_results', [False, True])
def test_golden_postprocess(name, async_results, cuda_device):
    z = load_golden('post_' + name)
    cfg = jload(z['cfg'])
    _, _, pan = _build(cfg, z['is_thing'], z['has_orientation'], async_results=async_results)
    r = _run(pan, z['logits'], z['heat'], z['offset'], z.get('orientation'), cuda_device)
    g = lambda k: r[k].cpu().numpy()
    assert np.array_equal(g('semantic_segmentation_idx'), z['semantic_segmentation_idx'])
    assert r['semantic_segmentation_idx'].dtype == torch.int64
    assert np.array_equal(g('panoptic_foreground_mask'), z['panoptic_foreground_mask'])
    assert np.array_equal(g('panoptic_segmentation_deeplab'), z['panoptic_segmentation_deeplab'])
    assert r['panoptic_segmentation_deeplab'].dtype == torch.int64
    assert np.array_equal(g('panoptic_segmentation_deeplab_instance_idx'),
                          z['panoptic_segmentation_deeplab_instance_idx'])
    assert r['panoptic_segmentation_deeplab_instance_idx'].dtype == torch.uint8
    assert np.array_equal(g('panoptic_segmentation_deeplab_semantic_idx'),
                          z['panoptic_segmentation_deeplab_semantic_idx'])
    assert r['panoptic_segmentation_deeplab_ids'] == int_keys(jload(z['ids']))
    _check_meta(jload(z['meta']), r['panoptic_segmentation_deeplab_instance_meta'])
    np.testing.assert_allclose(g('semantic_segmentation_score'), z['semantic_segmentation_score'],
                               rtol=1e-5)
    if 'orientations' in z:
        _check_orient(int_keys(jload(z['orientations'])),
                      r['orientations_panoptic_segmentation_deeplab_instance'])
    if 'semantic_score' in z:
        for k in ('semantic_score', 'instance_score', 'panoptic_score'):
            np.testing.assert_allclose(g(f'panoptic_segmentation_deeplab_{k}'), z[k], rtol=1e-5,
                                       atol=1e-7)
    # identity full-res twins alias the same results
    assert np.array_equal(r['panoptic_segmentation_deeplab_fullres'].cpu().numpy(),
                          z['panoptic_segmentation_deeplab'])


@pytest.mark.parametrize('async_results', [False, True])
def test_golden_postprocess_wrapped_instance_ids(async_results, cuda_device):
    """`post_wrap` (from the unmodified reference): two of three frames have 432 tied centres.
    Default: the error is explicit.  `on_overflow='wrap'`: the frames are redone with all centres
    and ids mod 256 -- maps, id dicts, the 432-entry meta dicts (zero areas beyond 255, NaN
    orientations) and the orientations equal the reference's (instance.py:231-266)."""
    from nicr_mt_scene_analysis_b200._lib import ERR_TOO_MANY_CENTERS, NpbError
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    z = load_golden('post_wrap')
    cfg = jload(z['cfg'])
    _, _, pan = _build(cfg, z['is_thing'], z['has_orientation'], async_results=async_results)
    with pytest.raises(NpbError) as err:
        r = _run(pan, z['logits'], z['heat'], z['offset'], z.get('orientation'), cuda_device)
        r['panoptic_segmentation_deeplab_ids']
    assert err.value.code == ERR_TOO_MANY_CENTERS

    ins = get_postprocessing_class(
        'instance', heatmap_threshold=cfg['thr'], heatmap_nms_kernel_size=cfg['ks'],
        top_k_instances=cfg['top_k'], heatmap_apply_foreground_mask=cfg['apply_fg'],
        normalized_offset=cfg['normalized'], offset_distance_threshold=cfg['dist_thr'],
        on_overflow='wrap')()
    pan = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=ins,
        semantic_classes_is_thing=tuple(bool(x) for x in z['is_thing']),
        semantic_class_has_orientation=tuple(bool(x) for x in z['has_orientation']),
        normalized_offset=cfg['normalized'], async_results=async_results)()
    for _ in range(2):      # (the second call reuses the workspace of the first)
        r = _run(pan, z['logits'], z['heat'], z['offset'], z.get('orientation'), cuda_device)
        g = lambda k: r[k].cpu().numpy()
        assert np.array_equal(g('panoptic_segmentation_deeplab_instance_idx'),
                              z['panoptic_segmentation_deeplab_instance_idx'])
        assert np.array_equal(g('panoptic_segmentation_deeplab'), z['panoptic_segmentation_deeplab'])
        assert np.array_equal(g('panoptic_segmentation_deeplab_semantic_idx'),
                              z['panoptic_segmentation_deeplab_semantic_idx'])
        assert r['panoptic_segmentation_deeplab_ids'] == int_keys(jload(z['ids']))
        meta = r['panoptic_segmentation_deeplab_instance_meta']
        assert [len(m) for m in meta] == [5, 432, 432]
        _check_meta(jload(z['meta']), meta)
        _check_orient(int_keys(jload(z['orientations'])),
                      r['orientations_panoptic_segmentation_deeplab_instance'])
    # the stage functions of the instance post-processing (ground-truth foreground path)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    fg = z['panoptic_foreground_mask']
    mask, centers = ins._get_instance_centers(t(z['heat']), t(fg))
    assert np.array_equal(mask.cpu().numpy(), z['center_mask'])
    assert [c.tolist() for c in centers] == jload(z['centers'])
    H, W = fg.shape[-2:]
    off_px = z['offset'] * np.array([H, W], np.float32).reshape(1, 2, 1, 1)
    seg, meta = ins._get_instance_segmentation(t(z['heat']), t(off_px.astype(np.float32)), t(fg))
    with oracle.allow_wrap():
        o_seg, o_meta = oracle.instance_segmentation(
            z['heat'], off_px.astype(np.float32), fg, cfg['thr'], cfg['ks'], cfg['top_k'],
            cfg['apply_fg'], False, cfg['dist_thr'], cap=1024)
    assert np.array_equal(seg.cpu().numpy(), o_seg)
    assert [len(m) for m in meta] == [5, 432, 432]
    for gm, om in zip(meta, o_meta):
        assert {k: (tuple(v['center_yx']), v['area']) for k, v in gm.items()} == \
            {k: (tuple(v['center_yx']), v['area']) for k, v in om.items()}
    # the fused evaluation counts a frame before it could be redone: refused up front
    with pytest.raises(ValueError):
        pan.fuse_evaluation(object())


def test_optional_stuff_area_filter(cuda_device):
    """`stuff_area=n` (Panoptic-DeepLab's filter; the reference has none, default off): stuff
    segments (`pan > 0 and pan % L == 0`) of fewer than n pixels become void, everything else --
    thing instances, id dicts, meta, the raw instance map -- stays as without the filter.  Checked
    against a numpy restatement of the rule on the golden case."""
    z = load_golden('post_q10')
    cfg = jload(z['cfg'])
    L = 1 << 16
    _, _, plain = _build(cfg, z['is_thing'], z['has_orientation'])
    r0 = _run(plain, z['logits'], z['heat'], z['offset'], z.get('orientation'), cuda_device)
    base = r0['panoptic_segmentation_deeplab'].cpu().numpy()
    assert np.array_equal(base, z['panoptic_segmentation_deeplab'])
    sizes = []
    for b in range(base.shape[0]):
        ids, cnt = np.unique(base[b], return_counts=True)
        sizes += [int(c) for i, c in zip(ids, cnt) if i > 0 and i % L == 0]
    limit = int(np.median(sizes)) + 1           # about half of the stuff segments go
    _, _, filt = _build(cfg, z['is_thing'], z['has_orientation'], stuff_area=limit)
    r1 = _run(filt, z['logits'], z['heat'], z['offset'], z.get('orientation'), cuda_device)
    want = base.copy()
    dropped = 0
    for b in range(base.shape[0]):
        ids, cnt = np.unique(base[b], return_counts=True)
        for i, c in zip(ids, cnt):
            if i > 0 and i % L == 0 and c < limit:
                want[b][base[b] == i] = 0
                dropped += 1
    assert 0 < dropped < len(sizes)
    got = r1['panoptic_segmentation_deeplab'].cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(r1['panoptic_segmentation_deeplab_semantic_idx'].cpu().numpy(), want // L)
    assert r1['panoptic_segmentation_deeplab_ids'] == r0['panoptic_segmentation_deeplab_ids']
    assert torch.equal(r1['panoptic_segmentation_deeplab_instance_idx'],
                       r0['panoptic_segmentation_deeplab_instance_idx'])
    assert torch.equal(r1['semantic_segmentation_idx'], r0['semantic_segmentation_idx'])
    with pytest.raises(ValueError):
        filt.fuse_evaluation(object())


def test_result_keys_like_reference(cuda_device):
    """tests/test_decoders+postprocessing.py:208-250 of the reference: key presence"""
    z = load_golden('post_scores')
    cfg = jload(z['cfg'])
    _, _, pan = _build(cfg, z['is_thing'], z['has_orientation'])
    r = _run(pan, z['logits'], z['heat'], z['offset'], z['orientation'], cuda_device)
    keys = ['semantic_output', 'semantic_side_outputs', 'semantic_softmax_scores',
            'semantic_segmentation_score', 'semantic_segmentation_idx', 'semantic_output_fullres',
            'semantic_softmax_scores_fullres', 'semantic_segmentation_score_fullres',
            'semantic_segmentation_idx_fullres', 'instance_output', 'instance_side_outputs',
            'instance_centers', 'instance_offsets', 'instance_orientation',
            'panoptic_foreground_mask', 'panoptic_segmentation_deeplab',
            'panoptic_segmentation_deeplab_fullres', 'panoptic_segmentation_deeplab_ids',
            'panoptic_segmentation_deeplab_semantic_idx',
            'panoptic_segmentation_deeplab_semantic_idx_fullres',
            'panoptic_segmentation_deeplab_semantic_score',
            'panoptic_segmentation_deeplab_semantic_score_fullres',
            'panoptic_segmentation_deeplab_instance_idx',
            'panoptic_segmentation_deeplab_instance_idx_fullres',
            'panoptic_segmentation_deeplab_instance_meta',
            'panoptic_segmentation_deeplab_instance_score',
            'panoptic_segmentation_deeplab_instance_score_fullres',
            'panoptic_segmentation_deeplab_panoptic_score',
            'panoptic_segmentation_deeplab_panoptic_score_fullres',
            'orientations_panoptic_segmentation_deeplab_instance']
    for k in keys:
        assert k in list(r.keys()), k
    r.materialize()
    assert r['semantic_softmax_scores'].shape == z['logits'].shape
    np.testing.assert_allclose(r['semantic_softmax_scores'].cpu().numpy(),
                               torch.softmax(torch.from_numpy(z['logits']), dim=1).numpy(),
                               rtol=1e-5, atol=1e-8)


def test_centers_tie_cases(cuda_device):
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    for c in jload(load_golden('centers')['cases']):
        post = get_postprocessing_class(
            'instance', heatmap_threshold=c['thr'], heatmap_nms_kernel_size=c['ks'],
            top_k_instances=c['k'], heatmap_apply_foreground_mask=c['apply_fg'])()
        heat = torch.tensor(c['heat'], dtype=torch.float32, device=cuda_device)
        fg = torch.tensor(c['fg'], dtype=torch.bool, device=cuda_device)
        mask, centers = post._get_instance_centers(heat, fg)
        assert np.array_equal(mask.cpu().numpy(), np.array(c['mask'], bool)), (c['ks'], c['k'])
        assert [x.tolist() for x in centers] == c['centers']
        assert all(x.dtype == torch.int32 for x in centers)


@pytest.mark.parametrize('cfg', [
    dict(B=3, C=40, H=120, W=160, K=12, seed=11, quantize='q10'),
    dict(B=2, C=37, H=106, W=146, K=20, seed=12, quantize='q10'),       # 530x730 / 5
    dict(B=2, C=19, H=128, W=256, K=100, seed=13, quantize='q10', top_k=100),
    dict(B=2, C=9, H=97, W=131, K=7, seed=14, quantize='tie', top_k=4),  # P % 4 != 0 -> scalar path
    dict(B=2, C=5, H=64, W=64, K=0, seed=15, quantize='q10'),            # no centres at all
    dict(B=1, C=1, H=32, W=48, K=2, seed=16, quantize='q10'),            # single class
    dict(B=2, C=255, H=24, W=32, K=3, seed=17, quantize='q10'),          # maximum class count
])
def test_against_oracle(cfg, cuda_device):
    from nicr_mt_scene_analysis_b200 import testing
    B, C, H, W, K = (cfg[k] for k in 'BCHWK')
    data = testing.make_batch(B, C, H, W, max(K, 1), seed=cfg['seed'], quantize=cfg['quantize'])
    if K == 0:
        data['heat'].zero_()
    is_thing = testing.default_is_thing(C) if C > 1 else (True,)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing)) if C > 1 else (True,)
    pcfg = dict(thr=0.1, ks=3, top_k=cfg.get('top_k', 64), apply_fg=False, normalized=True,
                dist_thr=None)
    _, _, pan = _build(pcfg, is_thing, has_ori)
    r = _run(pan, *(data[k].numpy() for k in ('logits', 'heat', 'offset', 'orientation')),
             cuda_device)
    ref = oracle.panoptic_postprocess(
        data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(),
        data['orientation'].numpy(), is_thing, has_ori, top_k=pcfg['top_k'])
    assert np.array_equal(r['_semantic_segmentation_idx_u8'].cpu().numpy(), ref['semantic_idx'])
    assert np.array_equal(r['panoptic_segmentation_deeplab_instance_idx'].cpu().numpy(),
                          ref['instance_idx'])
    assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
    assert r['panoptic_segmentation_deeplab_ids'] == ref['ids']
    for gm, rm in zip(r['panoptic_segmentation_deeplab_instance_meta'], ref['meta']):
        assert sorted(gm) == sorted(rm)
        for i in rm:
            assert gm[i]['center_yx'] == rm[i]['center_yx'] and gm[i]['area'] == rm[i]['area']
    _check_orient(ref['orientations'], r['orientations_panoptic_segmentation_deeplab_instance'])


@pytest.mark.parametrize('mode', ['dense_ties', 'wild_offsets', 'nonfinite_offsets'])
def test_grouping_with_many_centres_is_exact(mode, cuda_device):
    """stress for the per-warp centre pruning of group_pixels_kernel: up to 254 centres, exact
    distance ties (integer pixel targets), offsets pointing all over the frame, NaN / Inf"""
    from nicr_mt_scene_analysis_b200.model.postprocessing import InstancePostprocessing
    g = torch.Generator().manual_seed({'dense_ties': 1, 'wild_offsets': 2, 'nonfinite_offsets': 3}[mode])
    B, H, W = 2, 120, 200
    heat = torch.zeros(B, 1, H, W)
    ys = torch.arange(4, H - 4, 7)
    xs = torch.arange(4, W - 4, 13)
    heat[:, 0, ys[:, None], xs[None, :]] = 0.5 + 0.5 * torch.rand(B, len(ys), len(xs), generator=g)
    n_planted = len(ys) * len(xs)
    assert 200 < n_planted <= 254
    fg = torch.rand(B, H, W, generator=g) > 0.2
    if mode == 'dense_ties':          # every pixel points exactly at an integer location
        off = torch.stack((torch.randint(-9, 10, (B, H, W), generator=g).float(),
                           torch.randint(-9, 10, (B, H, W), generator=g).float()), 1)
    else:
        off = torch.stack(((torch.rand(B, H, W, generator=g) - 0.5) * 2 * H,
                           (torch.rand(B, H, W, generator=g) - 0.5) * 2 * W), 1)
    if mode == 'nonfinite_offsets':
        bad = torch.rand(B, 2, H, W, generator=g)
        off[bad < 0.02] = float('nan')
        off[(bad > 0.02) & (bad < 0.04)] = float('inf')
        off[(bad > 0.04) & (bad < 0.06)] = float('-inf')
    post = InstancePostprocessing(normalized_offset=False, top_k_instances=254)
    seg, meta = post._get_instance_segmentation(heat.to(cuda_device), off.to(cuda_device),
                                                fg.to(cuda_device))
    ref_seg, ref_meta = oracle.instance_segmentation(heat.numpy(), off.numpy(), fg.numpy(),
                                                     top_k=254, normalized_offset=False)
    assert [len(m) for m in ref_meta] == [n_planted] * B
    assert np.array_equal(seg.cpu().numpy(), ref_seg)
    assert [{k: v['area'] for k, v in m.items()} for m in meta] == \
        [{k: v['area'] for k, v in m.items()} for m in ref_meta]


def test_too_many_centers_raises(cuda_device):
    """> 255 centres (only reachable through k-th value ties): the reference wraps uint8
    ids silently (instance.py:236); this implementation refuses."""
    from nicr_mt_scene_analysis_b200 import _lib
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    heat = torch.zeros(1, 1, 96, 96, device=cuda_device)
    heat[:, :, 2:94:3, 2:94:3] = 0.5          # 31 * 31 = 961 equal isolated peaks
    post = get_postprocessing_class('instance', top_k_instances=10)()
    with pytest.raises(_lib.NpbError) as e:
        post._get_instance_centers(heat)
    assert e.value.code == _lib.ERR_TOO_MANY_CENTERS


def test_instance_postprocessing_like_reference_test(cuda_device):
    """port of the reference's tests/test_instance_postprocessing.py:90-150 (rectangles with
    exact pixel offsets; found centres == planted centres; segmentation is a relabelling)"""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import InstancePostprocessing
    g = torch.Generator().manual_seed(3)
    B, H, W = 4, 480, 640
    inst = torch.zeros(B, 1, H, W, dtype=torch.uint8)
    heat = torch.zeros(B, 1, H, W)
    off = torch.zeros(B, 2, H, W)
    fg = torch.zeros(B, 1, H, W, dtype=torch.bool)
    yy = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    xx = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
    planted = []
    for b in range(B):
        cs = []
        for i in range(2):
            x = int(torch.randint(0, W, (1,), generator=g))
            y = int(torch.randint(0, H, (1,), generator=g))
            s = int(torch.randint(20, 40, (1,), generator=g))
            x0, x1, y0, y1 = max(x - s, 0), min(x + s, W), max(y - s, 0), min(y + s, H)
            cx, cy = int(x1 - (x1 - x0) / 2), int(y1 - (y1 - y0) / 2)
            heat[b, 0, cy, cx] = 1
            cs.append((cy, cx))
            inst[b, 0, y0:y1, x0:x1] = i + 1
            off[b, 0, y0:y1, x0:x1] = cy - yy[y0:y1, x0:x1]
            off[b, 1, y0:y1, x0:x1] = cx - xx[y0:y1, x0:x1]
            fg[b, 0, y0:y1, x0:x1] = True
        planted.append(sorted(cs))
    post = InstancePostprocessing(normalized_offset=False)
    _, found = post._get_instance_centers(heat.to(cuda_device))
    _, ref_centers = oracle.instance_centers(heat.numpy())
    assert [f.tolist() for f in found] == [c.tolist() for c in ref_centers]
    for f, cs in zip(found, planted):       # every planted centre is found (ref test :106-112)
        assert set(map(tuple, f.tolist())) == set(cs)
    batch = testing.make_batch_dict(B, H, W)
    batch['instance_foreground'] = fg
    r = post.postprocess(((heat.to(cuda_device), off.to(cuda_device)), None), batch,
                         is_training=False)
    seg = r['instance_segmentation_gt_foreground'].cpu()
    assert 'instance_segmentation_gt_foreground_fullres' in r and 'instance_segmentation_gt_meta' in r
    ref_seg, ref_meta = oracle.instance_segmentation(heat.numpy(), off.numpy(), fg[:, 0].numpy(),
                                                     normalized_offset=False)
    assert np.array_equal(seg.numpy(), ref_seg)
    for gm, rm in zip(r['instance_segmentation_gt_meta'], ref_meta):
        assert {k: (v['center_yx'], v['area']) for k, v in gm.items()} == \
            {k: (v['center_yx'], v['area']) for k, v in rm.items()}


def test_merge_standalone_golden(cuda_device):
    from nicr_mt_scene_analysis_b200.utils import deeplab_merge_batch
    z = load_golden('merge')
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    pan, ids = deeplab_merge_batch(t(z['sem']), t(z['ins']), t(z['fg']), int(z['L']),
                                   z['thing_ids'], 0)
    assert np.array_equal(pan.cpu().numpy(), z['pan'])
    assert ids == int_keys(jload(z['ids']))
    pan, ids = deeplab_merge_batch(t(z['sem']), t(z['ins']), t(z['fg']), 1000, z['thing_ids'], 3)
    assert np.array_equal(pan.cpu().numpy(), z['pan_L1000_void3'])
    assert ids == int_keys(jload(z['ids_L1000_void3']))


def test_orientation_standalone_golden(cuda_device):
    from nicr_mt_scene_analysis_b200.model.postprocessing import InstancePostprocessing
    z = load_golden('orientation')
    post = InstancePostprocessing()
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    _check_orient(int_keys(jload(z['with_mask'])),
                  post._get_instance_orientation(t(z['ori']), t(z['seg']), t(z['mask'])))
    _check_orient(int_keys(jload(z['without_mask'])),
                  post._get_instance_orientation(t(z['ori']), t(z['seg']), None))


def test_cpu_tensors_are_rejected(cuda_device):
    """no CPU fallback: host tensors raise instead of silently running somewhere else"""
    z = load_golden('post_q10')
    cfg = jload(z['cfg'])
    _, _, pan = _build(cfg, z['is_thing'], z['has_orientation'])
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        _run(pan, z['logits'], z['heat'], z['offset'], z['orientation'], torch.device('cpu'))


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float64])
def test_half_precision_decoder_outputs_are_widened(dtype, cuda_device):
    """Decoder outputs of a network that ran under autocast: the kernels compute in float32, and
    widening is exact -- every result equals the float32 path on the same (rounded) values."""
    from nicr_mt_scene_analysis_b200 import testing
    z = load_golden('post_q10')
    cfg = jload(z['cfg'])
    _, _, pan = _build(cfg, z['is_thing'], z['has_orientation'])
    B, _, H, W = z['logits'].shape
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device).to(dtype)
    raw = (t(z['logits']), (t(z['heat']), t(z['offset']), t(z['orientation'])))
    wide = (raw[0].float(), tuple(x.float() for x in raw[1]))
    batch = testing.make_batch_dict(B, H, W)
    got = pan.postprocess((raw, (None, None)), batch, is_training=False)
    want = pan.postprocess((wide, (None, None)), batch, is_training=False)
    for k in ('semantic_segmentation_idx', 'panoptic_segmentation_deeplab',
              'panoptic_segmentation_deeplab_instance_idx'):
        assert torch.equal(got[k], want[k]), k
    assert got['panoptic_segmentation_deeplab_ids'] == want['panoptic_segmentation_deeplab_ids']
    _check_meta(want['panoptic_segmentation_deeplab_instance_meta'],
                got['panoptic_segmentation_deeplab_instance_meta'])
    assert torch.equal(got['semantic_segmentation_score'], want['semantic_segmentation_score'])


def test_full_size_properties(cuda_device):
    """BASELINE-size frame (530x730, C=37, 20 centres): size-independent invariants +
    agreement with the oracle on one frame."""
    from nicr_mt_scene_analysis_b200 import testing
    B, C, H, W, K = 2, 37, 530, 730, 20
    data = testing.make_batch(B, C, H, W, K, seed=5)
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    pcfg = dict(thr=0.1, ks=3, top_k=64, apply_fg=False, normalized=True, dist_thr=None)
    _, _, pan = _build(pcfg, is_thing, has_ori)
    r = _run(pan, *(data[k].numpy() for k in ('logits', 'heat', 'offset', 'orientation')),
             cuda_device)
    pan_map = r['panoptic_segmentation_deeplab']
    sem = r['semantic_segmentation_idx']
    inst = r['panoptic_segmentation_deeplab_instance_idx']
    L = 1 << 16
    thing = torch.tensor(is_thing, device=cuda_device)
    # instances only on thing pixels; stuff pixels carry (class+1)*L; ids decode consistently
    assert bool(((inst > 0) <= thing[sem]).all())
    stuff = ~thing[sem]
    assert bool((pan_map[stuff] == (sem[stuff] + 1) * L).all())
    assert bool((pan_map[thing[sem] & (inst == 0)] == 0).all())
    for b in range(B):
        for pan_id, ins_id in r['panoptic_segmentation_deeplab_ids'][b].items():
            assert bool(((pan_map[b] == pan_id) == (inst[b] == ins_id)).all())
        areas = {i: m['area'] for i, m in r['panoptic_segmentation_deeplab_instance_meta'][b].items()}
        counts = torch.bincount(inst[b].flatten().long(), minlength=256).cpu()
        assert all(int(counts[i]) == a for i, a in areas.items())
    ref = oracle.panoptic_postprocess(data['logits'].numpy()[:1], data['heat'].numpy()[:1],
                                      data['offset'].numpy()[:1], data['orientation'].numpy()[:1],
                                      is_thing, has_ori)
    assert np.array_equal(pan_map[:1].cpu().numpy(), ref['panoptic'])


def test_fullres_golden(cuda_device):
    """network resolution != dataset resolution (dense_base.py:15-58): nearest-resized index
    maps bit-exact; bilinear-resized logits to f32 rounding; the full-res class map may only
    differ from the reference where the two best resized logits are (nearly) tied."""
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    z = load_golden('fullres')
    B, C, H, W = z['logits'].shape
    FH, FW = (int(v) for v in z['fullres_shape'])
    pan = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=tuple(bool(x) for x in z['is_thing']),
        semantic_class_has_orientation=tuple(bool(x) for x in z['has_orientation']))()
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    batch = {'semantic_fullres': torch.zeros(B, FH, FW), 'instance_fullres': torch.zeros(B, FH, FW),
             '_applied_preprocessing': [[{'type': 'Resize',
                                          'valid_region_slice_y': slice(*z['valid_y'].tolist()),
                                          'valid_region_slice_x': slice(*z['valid_x'].tolist())}]] * B}
    r = pan.postprocess(((t(z['logits']), (t(z['heat']), t(z['offset']), t(z['orientation']))),
                         (None, None)), batch, is_training=False)
    for k in ('panoptic_segmentation_deeplab_fullres',
              'panoptic_segmentation_deeplab_instance_idx_fullres',
              'panoptic_segmentation_deeplab_semantic_idx_fullres'):
        assert r[k].shape[-2:] == (FH, FW)
        assert np.array_equal(r[k].cpu().numpy(), z[k]), k
    full = r['semantic_output_fullres'].cpu().numpy()
    np.testing.assert_allclose(full, z['semantic_output_fullres'], rtol=1e-5, atol=2e-6)
    idx = r['semantic_segmentation_idx_fullres'].cpu().numpy()
    differs = idx != z['semantic_segmentation_idx_fullres']
    top2 = np.sort(z['semantic_output_fullres'], axis=1)[:, -2:]
    assert np.all((top2[:, 1] - top2[:, 0])[differs] < 1e-5)
    assert differs.mean() < 1e-3
    np.testing.assert_allclose(r['semantic_segmentation_score_fullres'].cpu().numpy()[~differs],
                               z['semantic_segmentation_score_fullres'][~differs], rtol=2e-5)


def test_naive_merge_ground_truth_targets(cuda_device):
    """PanopticTargetGenerator's merge (panoptic_merge.py:43-107) on the GPU: golden + oracle"""
    from nicr_mt_scene_analysis_b200.utils import naive_merge_semantic_and_instance_batch
    z = load_golden('naive_merge')
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    pan, ids = naive_merge_semantic_and_instance_batch(t(z['sem']), t(z['ins']), 1 << 16,
                                                       z['thing_ids'], 0)
    assert np.array_equal(pan.cpu().numpy(), z['pan'])
    assert ids == int_keys(jload(z['ids']))
    g = torch.Generator().manual_seed(5)
    sem = torch.randint(0, 30, (2, 48, 64), generator=g).repeat_interleave(4, 1) \
        .repeat_interleave(2, 2).to(torch.uint8)
    ins = torch.randint(0, 100, (2, 24, 32), generator=g).repeat_interleave(8, 1).repeat_interleave(4, 2)
    ins = (ins * 211) % 65536
    pan, ids = naive_merge_semantic_and_instance_batch(sem.to(cuda_device), ins.to(cuda_device),
                                                       1000, [1, 4, 9, 29], 7)
    ref_pan, ref_ids = oracle.naive_merge_batch(sem.numpy(), ins.numpy(), 1000, [1, 4, 9, 29], 7)
    assert np.array_equal(pan.cpu().numpy(), ref_pan)
    assert ids == ref_ids


def test_instance_target_generator(cuda_device):
    """InstanceTargetGenerator (data/preprocessing/instance.py:97-286), batched on the GPU"""
    from nicr_mt_scene_analysis_b200.utils import InstanceTargetGenerator
    z = load_golden('instance_targets')
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    gen = InstanceTargetGenerator(sigma=5, semantic_classes_is_thing=z['is_thing'].tolist())
    r = gen(t(z['sem']), t(z['ins']))
    assert np.array_equal(r['instance_center'].cpu().numpy(), z['instance_center'])   # bit-exact
    assert np.array_equal(r['instance_offset'].cpu().numpy(),
                          z['instance_offset'].transpose(0, 3, 1, 2))
    assert np.array_equal(r['instance_foreground'].cpu().numpy(), z['instance_foreground'])
    assert np.array_equal(r['instance_center_mask'].cpu().numpy(), z['instance_center_mask'])
    assert r['encoded_instances'] == jload(z['encoded'])
    assert r['skipped_instances_due_to_stuff'] == [[], [], []]
    gen_px = InstanceTargetGenerator(sigma=5, semantic_classes_is_thing=z['is_thing'].tolist(),
                                     normalized_offset=False)
    r = gen_px(t(z['sem']), t(z['ins']))
    assert r['instance_offset'].dtype == torch.int16
    assert np.array_equal(r['instance_offset'].cpu().numpy(),
                          z['instance_offset_px'].transpose(0, 3, 1, 2))
    # a bigger random case against the oracle (many instances, ids up to 65535)
    g = torch.Generator().manual_seed(8)
    B, H, W = 2, 200, 300
    ins = torch.randint(0, 400, (B, H // 10, W // 10), generator=g).repeat_interleave(10, 1) \
        .repeat_interleave(10, 2)
    ins = (ins * 163) % 65536
    sem = torch.randint(1, 12, (B, H // 10, W // 10), generator=g).repeat_interleave(10, 1) \
        .repeat_interleave(10, 2).to(torch.uint8)
    is_thing = [False] + [bool(c % 2) for c in range(1, 12)]
    ins[~torch.tensor(is_thing)[sem.long()]] = 0
    gen = InstanceTargetGenerator(sigma=8, semantic_classes_is_thing=is_thing)
    r = gen(sem.to(cuda_device), ins.to(cuda_device))
    ref = oracle.instance_targets(sem.numpy(), ins.numpy(), 8, is_thing, True)
    for k in ref:
        assert np.array_equal(r[k].cpu().numpy(), ref[k]), k
    bad = ins.clone()
    bad[0][~torch.tensor(is_thing)[sem[0].long()]] = 7
    with pytest.raises(AssertionError):
        gen(sem.to(cuda_device), bad.to(cuda_device))


def test_semantic_postprocessing_standalone(cuda_device):
    """SemanticPostprocessing on its own (semantic.py:37-82): keys, dtypes, values"""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    z = load_golden('post_q10')
    B, C, H, W = z['logits'].shape
    post = get_postprocessing_class('semantic')()
    logits = torch.from_numpy(z['logits']).to(cuda_device)
    r = post.postprocess((logits, None), testing.make_batch_dict(B, H, W), is_training=False)
    for k in ('semantic_output', 'semantic_side_outputs', 'semantic_softmax_scores',
              'semantic_segmentation_score', 'semantic_segmentation_idx', 'semantic_output_fullres',
              'semantic_softmax_scores_fullres', 'semantic_segmentation_score_fullres',
              'semantic_segmentation_idx_fullres'):
        assert k in r
    assert r['semantic_segmentation_idx'].dtype == torch.int64
    assert np.array_equal(r['semantic_segmentation_idx'].cpu().numpy(),
                          z['semantic_segmentation_idx'])
    assert np.array_equal(r['semantic_segmentation_idx_fullres'].cpu().numpy(),
                          z['semantic_segmentation_idx'])
    np.testing.assert_allclose(r['semantic_segmentation_score'].cpu().numpy(),
                               z['semantic_segmentation_score'], rtol=1e-5)
    assert r['semantic_output_fullres'] is logits
    # training mode only forwards
    t = post.postprocess((logits, 'side'), {}, is_training=True)
    assert t == {'semantic_output': logits, 'semantic_side_outputs': 'side'}
    # out-of-scope tasks are refused loudly, unknown names like the reference
    with pytest.raises(NotImplementedError):
        get_postprocessing_class('normal')
    with pytest.raises(ValueError):
        get_postprocessing_class('nope')


@pytest.mark.parametrize('shape', [
    dict(name='nyuv2', B=8, C=40, H=480, W=640, K=12, top_k=64, ori=False),
    dict(name='scannet', B=1, C=40, H=968, W=1296, K=30, top_k=64, ori=False),
    dict(name='cityscapes', B=1, C=19, H=1024, W=2048, K=100, top_k=100, ori=False),
])
def test_baseline_shapes_full_size(shape, cuda_device):
    """the other BASELINE.json shapes at full frame size: bit-exact against the oracle,
    plus the evaluation checksum (every pixel lands in exactly one confusion cell / pair)"""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    B, C, H, W, K = (shape[k] for k in 'BCHWK')
    data = testing.make_batch(B, C, H, W, K, seed=77, with_orientation=False, quantize='q10')
    is_thing = testing.default_is_thing(C)
    pcfg = dict(thr=0.1, ks=3, top_k=shape['top_k'], apply_fg=False, normalized=True, dist_thr=None)
    _, _, pan = _build(pcfg, is_thing, (False,) * C)
    r = _run(pan, data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(), None,
             cuda_device)
    ref = oracle.panoptic_postprocess(data['logits'].numpy(), data['heat'].numpy(),
                                      data['offset'].numpy(), None, is_thing, None,
                                      top_k=shape['top_k'])
    assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
    assert np.array_equal(r['panoptic_segmentation_deeplab_instance_idx'].cpu().numpy(),
                          ref['instance_idx'])
    assert r['panoptic_segmentation_deeplab_ids'] == ref['ids']
    pred = r['panoptic_segmentation_deeplab']
    tgt, tgt_sem = testing.make_eval_targets(pred)
    pq = PanopticQuality(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=cuda_device)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=cuda_device)
    PanopticEvaluation(pq, miou).update(pred, tgt, tgt_sem)
    pq.check_status()
    assert int(miou.confmat.sum()) == B * H * W
    out = oracle.pq_compare_and_accumulate(ref['panoptic'][0], tgt[0].cpu().numpy(), C + 1, 0,
                                           1 << 16, 256 ** 3, 0)
    if B == 1:
        got = np.stack([getattr(pq, n).cpu().numpy() for n in
                        ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])
        assert np.array_equal(got, np.stack(out[:4]))


@pytest.mark.parametrize('shape', [(2, 40, 48, 64), (2, 37, 30, 50), (1, 9, 33, 47), (1, 1, 16, 20),
                                   (2, 8, 24, 32), (1, 13, 20, 28)])
def test_nonfinite_logits_against_oracle(shape, cuda_device):
    """NaN / +-Inf logits (semantic.py:52-53: softmax -> max answers class 0 with a NaN score as
    soon as a NaN or +Inf poisons the soft-max, -Inf next to finite logits is harmless): the fused
    arg-max of the grouping kernel (4-pixel and scalar path) and the stand-alone semantic
    post-processing against the oracle, which tests/test_oracle_live_reference.py pins on the
    live reference for the same poisoning."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    B, C, H, W = shape
    for frac in (0.03, 0.5):
        data = testing.make_batch(B, C, H, W, 3, seed=41, quantize='q10', with_orientation=False)
        testing.poison_logits(data['logits'], frac, seed=H + W)
        is_thing = tuple(bool(c % 2 == 0) for c in range(C))
        pcfg = dict(thr=0.1, ks=3, top_k=64, apply_fg=False, normalized=True, dist_thr=None)
        _, _, pan = _build(pcfg, is_thing, (False,) * C)
        r = _run(pan, data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(), None,
                 cuda_device)
        ref = oracle.panoptic_postprocess(data['logits'].numpy(), data['heat'].numpy(),
                                          data['offset'].numpy(), None, is_thing, (False,) * C)
        assert np.array_equal(r['_semantic_segmentation_idx_u8'].cpu().numpy(), ref['semantic_idx'])
        assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
        assert r['panoptic_segmentation_deeplab_ids'] == ref['ids']
        # score of the winner: NaN exactly where the reference's is, equal elsewhere
        want = oracle.semantic_score(data['logits'].numpy())
        np.testing.assert_allclose(r['semantic_segmentation_score'].cpu().numpy(), want, rtol=1e-5)
        sem = get_postprocessing_class('semantic')()
        rs = sem.postprocess((data['logits'].to(cuda_device), None), testing.make_batch_dict(B, H, W),
                             is_training=False)
        assert np.array_equal(rs['semantic_segmentation_idx'].cpu().numpy(), ref['semantic_idx'])
        np.testing.assert_allclose(rs['semantic_segmentation_score'].cpu().numpy(), want, rtol=1e-5)


def test_nonfinite_logits_through_the_fullres_resize(cuda_device):
    """the full-resolution class map is the arg-max of the soft-max of the RESIZED logits
    (semantic.py:68-74): a non-finite tap poisons every output pixel it contributes to."""
    import torch.nn.functional as F
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    B, C, H, W, FH, FW = 2, 7, 24, 32, 37, 53
    data = testing.make_batch(B, C, H, W, 2, seed=43, quantize='q10', with_orientation=False)
    testing.poison_logits(data['logits'], 0.03, seed=5)
    sem = get_postprocessing_class('semantic')()
    batch = {'semantic_fullres': torch.zeros(B, FH, FW),
             '_applied_preprocessing': [[{'type': 'Resize', 'valid_region_slice_y': slice(0, H),
                                          'valid_region_slice_x': slice(0, W)}]] * B}
    r = sem.postprocess((data['logits'].to(cuda_device), None), batch, is_training=False)
    full = F.interpolate(data['logits'], size=(FH, FW), mode='bilinear', align_corners=False)
    score, idx = torch.max(F.softmax(full, dim=1), dim=1)
    got = r['semantic_segmentation_idx_fullres'].cpu()
    finite = torch.isfinite(full).all(1)
    assert (~finite).any()
    # pixels touched by a non-finite tap: exactly the reference's rule (class 0 when poisoned)
    assert torch.equal(got[~finite], idx[~finite])
    # elsewhere the usual near-tie caveat of the resized logits applies
    top2 = torch.sort(full, dim=1)[0][:, -2:]
    differs = (got != idx) & finite
    assert bool(((top2[:, 1] - top2[:, 0])[differs] < 1e-5).all())
    got_score = r['semantic_segmentation_score_fullres'].cpu()
    assert torch.equal(torch.isnan(got_score), torch.isnan(score))


def test_unquantised_bench_frames(cuda_device):
    """bench.py's inputs are NOT quantised: the CUDA path against the oracle on exactly the frames
    of the bench's headline workload (seeds 1000 + i); the live-reference test
    test_unquantised_bench_inputs_have_no_softmax_flips shows the oracle == the reference's
    softmax -> max on the same 4.3 M pixels, so there is no hidden arg-max flip in the bench."""
    from nicr_mt_scene_analysis_b200 import testing
    C, H, W, K = 40, 480, 640, 12
    is_thing = testing.default_is_thing(C)
    pcfg = dict(thr=0.1, ks=3, top_k=64, apply_fg=False, normalized=True, dist_thr=None)
    _, _, pan = _build(pcfg, is_thing, (False,) * C)
    for lo in (0, 7):
        frames = [testing.make_frame(C, H, W, K, seed=1000 + i, with_orientation=False, quantize=None)
                  for i in range(lo, lo + 7)]
        data = {k: torch.stack([f[k] for f in frames]) for k in frames[0]}
        r = _run(pan, data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(), None,
                 cuda_device)
        ref = oracle.panoptic_postprocess(data['logits'].numpy(), data['heat'].numpy(),
                                          data['offset'].numpy(), None, is_thing, (False,) * C)
        assert np.array_equal(r['_semantic_segmentation_idx_u8'].cpu().numpy(), ref['semantic_idx'])
        assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
        assert r['panoptic_segmentation_deeplab_ids'] == ref['ids']

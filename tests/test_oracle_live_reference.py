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
        ks=int(rng.choice([3, 3, 5])), thr=float(rng.choice([0.1, 0.3])),
        apply_fg=bool(rng.integers(0, 2)), normalized=bool(rng.integers(0, 2)),
        dist_thr=(None if rng.integers(0, 2) else int(rng.integers(3, 25))),
        with_orientation=bool(rng.integers(0, 2)))


@pytest.mark.parametrize('seed', range(8))
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
    got = oracle.panoptic_postprocess(
        data['logits'].numpy(), data['heat'].numpy(), data['offset'].numpy(),
        data['orientation'].numpy() if c['with_orientation'] else None, is_thing, has_ori,
        threshold=c['thr'], nms_kernel_size=c['ks'], top_k=c['top_k'],
        apply_foreground_mask=c['apply_fg'], normalized_offset=c['normalized'],
        offset_distance_threshold=c['dist_thr'])
    assert np.array_equal(got['semantic_idx'], r['semantic_segmentation_idx'].numpy()), c
    assert np.array_equal(got['instance_idx'], r['panoptic_segmentation_deeplab_instance_idx'].numpy()), c
    assert np.array_equal(got['panoptic'], r['panoptic_segmentation_deeplab'].numpy()), c
    assert got['ids'] == [{int(k): int(v) for k, v in d.items()}
                          for d in r['panoptic_segmentation_deeplab_ids']], c
    for gm, rm in zip(got['meta'], r['panoptic_segmentation_deeplab_instance_meta']):
        assert {k: (tuple(v['center_yx']), v['area']) for k, v in gm.items()} == \
            {int(k): (tuple(int(x) for x in v['center_yx']), int(v['area'])) for k, v in rm.items()}, c
    if c['with_orientation']:
        for dg, dr in zip(got['orientations'], r['orientations_panoptic_segmentation_deeplab_instance']):
            assert sorted(dg) == sorted(int(k) for k in dr), c
            for k, v in dr.items():
                assert abs(dg[int(k)] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (c, k)

    # evaluation: the reference's compare_and_accumulate and confusion matrix on the same maps
    L, OFF = 1 << 16, 256 ** 3
    pred = r['panoptic_segmentation_deeplab']
    tgt = torch.roll(pred, int(rng.integers(1, 5)), dims=-1).contiguous()
    for b in range(B):
        try:
            want = ref['pq'](pred[b], tgt[b], C + 1, 0, L, OFF, 0)
        except ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), C + 1, 0, L, OFF, 0)
            continue
        have = oracle.pq_compare_and_accumulate(pred[b].numpy(), tgt[b].numpy(), C + 1, 0, L, OFF, 0)
        for w, h in zip(want[:4], have[:4]):
            assert np.array_equal(np.asarray(w, np.float64), h), c            # float64, bit for bit
        assert {(int(g), int(p)) for g, p in want[4]} == have[4], c
    m = ref['miou'](n_classes=C + 1, ignore_first_class=True)
    m.reset()
    m.update(preds=pred // L, target=(tgt // L).to(torch.uint8))
    assert np.array_equal(m.confmat.numpy(), oracle.confmat((pred // L).numpy(), (tgt // L).numpy(), C + 1)), c

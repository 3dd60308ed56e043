"""Host-buffer pipeline (the end-to-end path of bench.py) and CUDA-graph replay against the
oracle / the eager path."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _setup(C, dev, **kw):
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori, **kw)()
    pq = PanopticQuality(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    return post, PanopticEvaluation(pq, miou), is_thing, has_ori


def _oracle_eval(pan, tgt, tgt_sem, C):
    state = np.zeros((4, C + 1))
    for b in range(pan.shape[0]):
        out = oracle.pq_compare_and_accumulate(pan[b], tgt[b], C + 1, 0, 1 << 16, 256 ** 3, 0)
        for s, v in zip(state, out[:4]):
            s += v
    return state, oracle.confmat(pan // (1 << 16), tgt_sem, C + 1)


def test_host_pipeline_matches_oracle(cuda_device):
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.pipeline import PanopticHostPipeline
    B, C, H, W, K = 7, 11, 96, 132, 5          # 7 frames in chunks of 3: ragged last chunk
    post, ev, is_thing, has_ori = _setup(C, cuda_device, async_results=True)
    data = testing.make_batch(B, C, H, W, K, seed=21)
    ref = oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                        ('logits', 'heat', 'offset', 'orientation')),
                                      is_thing, has_ori)
    tgt = np.roll(ref['panoptic'], 5, axis=-1)
    tgt_sem = (tgt // (1 << 16)).astype(np.uint8)
    pinned = {k: v.pin_memory() for k, v in data.items()}
    targets = {'panoptic': torch.from_numpy(tgt).pin_memory(),
               'semantic': torch.from_numpy(tgt_sem).pin_memory()}
    pipe = PanopticHostPipeline(post, ev, chunk_frames=3, device=cuda_device)
    for _ in range(2):                            # second run reuses the staging slots
        ev.reset()
        out = PanopticHostPipeline.finish(pipe.run(pinned, testing.make_batch_dict(B, H, W), targets))
        assert not out['panoptic_segmentation_deeplab'].is_cuda
        assert np.array_equal(out['panoptic_segmentation_deeplab'].numpy(), ref['panoptic'])
        assert np.array_equal(out['panoptic_segmentation_deeplab_instance_idx'].numpy(),
                              ref['instance_idx'])
        assert out['panoptic_segmentation_deeplab_ids'] == ref['ids']
        assert [sorted(d) for d in out['orientations_panoptic_segmentation_deeplab_instance']] == \
            [sorted(d) for d in ref['orientations']]
        ev.pq.check_status()
        state, cm = _oracle_eval(ref['panoptic'], tgt, tgt_sem, C)
        got = np.stack([getattr(ev.pq, n).cpu().numpy() for n in
                        ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])
        # chunks are evaluated in frame order, so even the float64 sums are bit-identical
        assert np.array_equal(got, state)
        assert np.array_equal(ev.miou.confmat.cpu().numpy(), cm)
    assert pipe.h2d_bytes == sum(v.numel() * v.element_size() for v in pinned.values()) + \
        tgt.nbytes + tgt_sem.nbytes
    assert pipe.d2h_bytes >= B * H * W * 9


def test_graph_replay_equals_eager(cuda_device):
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    B, C, H, W, K = 4, 9, 64, 96, 4
    post, ev, is_thing, has_ori = _setup(C, cuda_device, async_results=True)
    bufs = {k: v.to(cuda_device) for k, v in testing.make_batch(B, C, H, W, K, seed=31).items()}
    batch = testing.make_batch_dict(B, H, W)
    tgt = torch.zeros((B, H, W), dtype=torch.int64, device=cuda_device)
    tgt_sem = torch.zeros((B, H, W), dtype=torch.uint8, device=cuda_device)

    def step():
        r = post.postprocess(((bufs['logits'], (bufs['heat'], bufs['offset'], bufs['orientation'])),
                              (None, None)), batch, is_training=False)
        ev.update(r['panoptic_segmentation_deeplab'], tgt, tgt_sem)
        return r

    captured = CapturedStep(step, warmup=2, device=cuda_device)
    for seed in (32, 33):            # new inputs are copied INTO the captured buffers
        fresh = testing.make_batch(B, C, H, W, K, seed=seed)
        ref = oracle.panoptic_postprocess(*(fresh[k].numpy() for k in
                                            ('logits', 'heat', 'offset', 'orientation')),
                                          is_thing, has_ori)
        t = np.roll(ref['panoptic'], 3, axis=-1)
        ts = (t // (1 << 16)).astype(np.uint8)
        for k in bufs:
            bufs[k].copy_(fresh[k])
        tgt.copy_(torch.from_numpy(t))
        tgt_sem.copy_(torch.from_numpy(ts))
        ev.reset()
        r = captured.replay()
        torch.cuda.synchronize()
        assert np.array_equal(r['panoptic_segmentation_deeplab'].cpu().numpy(), ref['panoptic'])
        assert r['_panoptic_instance_tables'].panoptic_ids() == ref['ids']
        state, cm = _oracle_eval(ref['panoptic'], t, ts, C)
        ev.pq.check_status()
        got = np.stack([getattr(ev.pq, n).cpu().numpy() for n in
                        ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')])
        assert np.array_equal(got, state)
        assert np.array_equal(ev.miou.confmat.cpu().numpy(), cm)


def test_host_pipeline_overlapped_batches(cuda_device):
    """Two batches enqueued back to back, finished afterwards (each with its own result
    buffers): `finish` waits for its own batch only and both results are right."""
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.pipeline import PanopticHostPipeline
    B, C, H, W, K = 5, 9, 80, 112, 4
    post, ev, is_thing, has_ori = _setup(C, cuda_device, async_results=True)
    pipe = PanopticHostPipeline(post, None, chunk_frames=2, device=cuda_device)
    batches, refs, pending = [], [], []
    for seed in (31, 32, 33):
        data = testing.make_batch(B, C, H, W, K, seed=seed)
        refs.append(oracle.panoptic_postprocess(*(data[k].numpy() for k in
                                                  ('logits', 'heat', 'offset', 'orientation')),
                                                is_thing, has_ori))
        batches.append({k: v.pin_memory() for k, v in data.items()})
    for pinned in batches:                       # enqueue everything first
        pending.append(pipe.run(pinned, testing.make_batch_dict(B, H, W)))
    for out, ref in zip(pending, refs):
        out = PanopticHostPipeline.finish(out)
        assert np.array_equal(out['panoptic_segmentation_deeplab'].numpy(), ref['panoptic'])
        assert np.array_equal(out['panoptic_segmentation_deeplab_instance_idx'].numpy(),
                              ref['instance_idx'])
        assert out['panoptic_segmentation_deeplab_ids'] == ref['ids']

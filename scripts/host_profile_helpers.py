#!/usr/bin/env python
"""Host time of the reference-facing calls around the kernels (synchronous `postprocess`,
`PanopticTaskHelper.validation_step`, `InstanceTaskHelper.validation_step`) on one batch of
the default shape, with cProfile: where python spends its time once the GPU part is ~1 ms."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    from nicr_mt_scene_analysis_b200.task_helper import InstanceTaskHelper, PanopticTaskHelper
    dev = torch.device('cuda', 0)
    w = bench.WORKLOAD
    B, C, H, W, K = w['B'], w['C'], w['H'], w['W'], w['K']
    L = 1 << 16
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    data = testing.make_batch(B, C, H, W, K, seed=3, with_orientation=True, device=dev, quantize=None)
    batch = testing.make_batch_dict(B, H, W)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori)()
    raw = ((data['logits'], (data['heat'], data['offset'], data['orientation'])), (None, None))

    def run_post():
        r = post.postprocess(raw, batch, is_training=False)
        r.materialize() if hasattr(r, 'materialize') else None
        return r

    r = run_post()
    pan = r['panoptic_segmentation_deeplab']
    tgt = torch.roll(pan, 5, -1).contiguous()
    ids = r['panoptic_segmentation_deeplab_ids']
    ori = r['orientations_panoptic_segmentation_deeplab_instance']
    tgt_ids = [{int(k): int(k) % L for k in torch.unique(tgt[b]).tolist() if int(k) % L} for b in range(B)]
    ori_t = [{i: 0.1 * i for i in set(d.values())} for d in tgt_ids]
    vbatch = dict(batch, panoptic_fullres=tgt, semantic_fullres=(tgt // L).to(torch.uint8),
                  instance_fullres=(tgt % L).to(torch.int32), panoptic_ids_to_instance_dict=tgt_ids,
                  orientations_present=ori_t)
    pan_helper = PanopticTaskHelper(C + 1, (False,) + is_thing)
    pan_helper.initialize(dev)
    ins_helper = InstanceTaskHelper(C + 1, (False,) + is_thing)
    ins_helper.initialize(dev)
    pan_post = {'panoptic_segmentation_deeplab_fullres': pan, 'panoptic_segmentation_deeplab_ids': ids,
                'orientations_panoptic_segmentation_deeplab_instance': ori}
    inst_fg = r['panoptic_segmentation_deeplab_instance_idx']
    ins_post = {'instance_segmentation_gt_foreground_fullres': inst_fg,
                'orientations_instance_segmentation_gt_orientation_foreground': [
                    {i: 0.2 * i for i in range(1, 256)} for _ in range(B)],
                'orientations_gt_instance_gt_orientation_foreground': ori_t}

    calls = [('postprocess (synchronous, all entries materialised)', run_post),
             ('PanopticTaskHelper.validation_step', lambda: pan_helper.validation_step(vbatch, 1, pan_post)),
             ('InstanceTaskHelper.validation_step', lambda: ins_helper.validation_step(vbatch, 1, ins_post))]
    for name, fn in calls:
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        print(f'== {name}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per {B}-frame call')
        prof = cProfile.Profile()
        prof.enable()
        fn()
        torch.cuda.synchronize(dev)
        prof.disable()
        pstats.Stats(prof).sort_stats('cumulative').print_stats(14)


if __name__ == '__main__':
    main()

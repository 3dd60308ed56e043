#!/usr/bin/env python
"""Achieved HBM bandwidth of the stand-alone entry points (the ones outside the fused panoptic
step): algorithmic bytes / CUDA-event time, as a fraction of the measured copy peak.

    python scripts/bench_entries.py [--frames 64 --height 530 --width 730 --classes 37]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=64)
    ap.add_argument('--height', type=int, default=530)
    ap.add_argument('--width', type=int, default=730)
    ap.add_argument('--classes', type=int, default=37)
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    import torch
    import bench
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    from nicr_mt_scene_analysis_b200.model.postprocessing.semantic import (
        semantic_argmax, semantic_argmax_resized, softmax_scores, widen_u8)
    from nicr_mt_scene_analysis_b200.utils.panoptic_merge import (
        deeplab_merge_batch, naive_merge_semantic_and_instance_batch)
    dev = torch.device('cuda', 0)
    B, H, W, C = args.frames, args.height, args.width, args.classes
    P = H * W
    K = 20
    pool = 8
    frames = [testing.make_frame(C, H, W, K, seed=50 + i, with_orientation=True, device=dev,
                                 quantize=None) for i in range(pool)]
    data = {k: torch.stack([frames[i % pool][k] for i in range(B)]).contiguous() for k in frames[0]}
    del frames
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori)()
    r = post.postprocess(((data['logits'], (data['heat'], data['offset'], data['orientation'])),
                          (None, None)), testing.make_batch_dict(B, H, W), is_training=False)
    sem_u8 = r['panoptic_segmentation_deeplab_semantic_idx'].to(torch.uint8).contiguous()   # 0 = void
    inst_u8 = r['panoptic_segmentation_deeplab_instance_idx'].contiguous()
    fg = r['panoptic_foreground_mask'].contiguous()
    sem_i64 = sem_u8.to(torch.int64)
    inst_i32 = inst_u8.to(torch.int32)
    thing_ids = [c + 1 for c, t in enumerate(is_thing) if t]
    peak, _ = bench.measured_peak_gbs()

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / args.reps

    h2, w2 = (H * 5) // 4, (W * 5) // 4
    entries = [
        ('semantic_argmax (classes only)', lambda: semantic_argmax(data['logits']), P * (4 * C + 1)),
        ('semantic_argmax + score', lambda: semantic_argmax(data['logits'], True), P * (4 * C + 5)),
        ('softmax_scores', lambda: softmax_scores(data['logits']), P * 8 * C),
        ('semantic_argmax_resized x1.25', lambda: semantic_argmax_resized(data['logits'], (0, 0, H, W), (h2, w2)),
         P * 4 * C + h2 * w2 * 5),
        ('widen_u8', lambda: widen_u8(sem_u8), P * 9),
        ('deeplab_merge_batch (int64 semantic, host dicts)',
         lambda: deeplab_merge_batch(sem_i64, inst_u8, fg, 1 << 16, thing_ids, 0), P * (8 + 1 + 1 + 8)),
        ('naive_merge_semantic_and_instance_batch',
         lambda: naive_merge_semantic_and_instance_batch(sem_u8, inst_i32, 1 << 16, thing_ids, 0), P * (1 + 4 + 8)),
    ]
    post_async = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
        async_results=True)()
    post_scores = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
        compute_scores=True, async_results=True)()
    raw = ((data['logits'], (data['heat'], data['offset'], data['orientation'])), (None, None))
    bdict = testing.make_batch_dict(B, H, W)
    bytes_post = bench.bytes_post_per_frame(C, H, W, True)
    entries += [
        ('postprocess (async results)', lambda: post_async.postprocess(raw, bdict, is_training=False),
         bytes_post),
        ('postprocess with compute_scores (async results; + logits again, 3 score maps)',
         lambda: post_scores.postprocess(raw, bdict, is_training=False),
         bytes_post + P * (4 * C + 2 + 12 + 4)),
    ]
    out = {'frames': B, 'height': H, 'width': W, 'classes': C, 'peak_GBs': peak}
    for name, fn, bytes_per_frame in entries:
        try:
            ms = timed(fn)
            gbs = bytes_per_frame * B / ms / 1e6
            out[name] = {'ms': round(ms, 4), 'GBs': round(gbs, 1), 'frac_of_peak': round(gbs / peak, 3)}
        except Exception as exc:      # keep the table going, say what happened
            out[name] = {'error': repr(exc)}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()

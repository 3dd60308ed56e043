#!/usr/bin/env python
"""BASELINE.json configs[4]: mIoU + PQ accumulation over 50 k synthetic 480x640 frames with the
confusion-matrix / PQ all-reduce at compute(), on 1..8 GPUs (torchrun).  Evaluation only:
17 algorithmic bytes per pixel (int64 prediction, int64 target, uint8 semantic target).

    python scripts/bench_eval.py [--frames 50000] [--batch 256] [--height 480 --width 640 --classes 40]
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_eval.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=50000)
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--classes', type=int, default=40)
    ap.add_argument('--instances', type=int, default=12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    C, H, W, K, B = args.classes, args.height, args.width, args.instances, args.batch
    L, OFF = 1 << 16, 256 ** 3
    is_thing = testing.default_is_thing(C)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=(False,) * C,
        async_results=True)()
    # a pool of distinct predicted frames (through the real post-processing), cycled to B
    pool = 32
    preds = []
    for i in range(0, pool, 8):
        d = testing.make_batch(8, C, H, W, K, seed=100 * (rank + 1) + i, with_orientation=False,
                               device=dev, quantize=None)
        r = post.postprocess(((d['logits'], (d['heat'], d['offset'])), (None, None)),
                             testing.make_batch_dict(8, H, W), is_training=False)
        preds.append(r['panoptic_segmentation_deeplab'].clone())
    preds = torch.cat(preds)
    pred = preds[torch.arange(B, device=dev) % pool].contiguous()
    tgt, tgt_sem = testing.make_eval_targets(pred, L)
    pq = PanopticQuality(C + 1, 0, L, OFF, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    ev = PanopticEvaluation(pq, miou)
    steps = -(-args.frames // (B * world))
    step = CapturedStep(lambda: ev.update(pred, tgt, tgt_sem), warmup=3, device=dev).replay
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        step()
    ev.compute()
    ev.reset()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    res = ev.compute()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    frames = steps * B * world
    peak, src = bench.measured_peak_gbs()
    if rank == 0:
        fps = frames / (ms * 1e-3)
        print(json.dumps({
            'metric': f'mIoU+PQ accumulation frames/s @{H}x{W}', 'value': fps, 'unit': 'frames/s',
            'n_gpus': world, 'frames': frames, 'steps': steps, 'batch_per_gpu': B, 'ms_total': ms,
            'roofline': {'bound': 'hbm', 'bytes_per_frame': 17 * H * W,
                         'achieved': fps / world * 17 * H * W / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': fps / world * 17 * H * W / 1e9 / peak, 'peak_source': src},
            'quality': {'all_pq': float(res['all_pq']), 'miou': float(res['semantic_miou'])}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

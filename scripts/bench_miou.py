#!/usr/bin/env python
"""Stand-alone mIoU update (`MeanIntersectionOverUnion.update` / `update_nonvoid`, the call of
SemanticTaskHelper.validation_step): int64 predictions + uint8 targets = 9 algorithmic bytes
per pixel.  Times the streaming kernel (aligned maps) and, for comparison, the generic
one-element-per-thread kernel (the same maps started one element later) with CUDA events.

    python scripts/bench_miou.py [--frames 256] [--height 480 --width 640 --classes 40]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=256)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--classes', type=int, default=40)
    ap.add_argument('--reps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=1000)
    args = ap.parse_args()
    import torch
    import bench
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion
    dev = torch.device('cuda', 0)
    B, H, W, C = args.frames, args.height, args.width, args.classes
    g = torch.Generator(device=dev).manual_seed(1)
    # blocky targets (0 = void), predictions = target - 1 with 8 % isolated wrong pixels
    low = torch.randint(0, C + 1, (B, (H + 31) // 32, (W + 31) // 32), generator=g, device=dev)
    target = low.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, :H, :W].contiguous()
    preds = torch.where(torch.rand(B, H, W, generator=g, device=dev) < 0.08,
                        torch.randint(0, C, (B, H, W), generator=g, device=dev),
                        (target - 1).clamp(min=0))
    N = B * H * W
    flat_p = torch.empty(N + 16, dtype=torch.int64, device=dev)
    flat_t = torch.empty(N + 16, dtype=torch.uint8, device=dev)
    peak, peak_src = bench.measured_peak_gbs()
    out = {'frames': B, 'height': H, 'width': W, 'classes': C, 'bytes_per_pixel': 9,
           'peak_GBs': peak, 'peak_source': peak_src}
    for name, first in (('streaming', 0), ('generic', 1)):
        p, t = flat_p[first:first + N], flat_t[first:first + N]
        p.copy_(preds.view(-1)); t.copy_(target.view(-1).to(torch.uint8))
        for entry in ('update_nonvoid', 'update'):
            m = MeanIntersectionOverUnion(n_classes=C + (entry == 'update'), device=dev)
            fn = getattr(m, entry)
            for _ in range(args.warmup):        # long enough for the clocks to settle
                fn(p, t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(args.reps):
                fn(p, t)
            e1.record()
            torch.cuda.synchronize(dev)
            m.check_status()
            ms = e0.elapsed_time(e1) / args.reps
            gbs = 9 * N / ms / 1e6
            out[f'{name}_{entry}'] = {'ms_per_launch': ms, 'frames_per_s': B / ms * 1e3,
                                      'GBs': gbs, 'frac_of_peak': gbs / peak}
            want = int(((t != 0) if entry == 'update_nonvoid' else torch.ones_like(t, dtype=torch.bool)).sum())
            assert int(m.confmat.sum()) == want * (args.warmup + args.reps)
    print(json.dumps(out))


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Per-step host time of the drop-in validation loop (bench.py `value_api`), step by step: where
does a slow start come from?   python scripts/probes/api_step_times.py [config] [fused]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from nicr_mt_scene_analysis_b200.task_helper import PanopticTaskHelper  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'sunrgbd'
    fused = len(sys.argv) > 2 and sys.argv[2] == 'fused'
    dev = torch.device('cuda', 0)
    w = dict(bench.WORKLOADS[cfg])
    arm = bench.Arm(w, w['B'], dev, 0, fused=False, graph=False)
    post = arm.new_post()
    helper = PanopticTaskHelper(w['C'] + 1, (False,) + arm.is_thing)
    helper.initialize(dev)
    if fused:
        post.fuse_evaluation(helper.evaluation)
    batch = dict(arm.batch, panoptic_fullres=arm.tgt_pan, semantic_fullres=arm.tgt_sem)
    import gc
    gc_log = []

    def on_gc(phase, info, _t=[0.0]):
        if phase == 'start':
            _t[0] = time.perf_counter()
        else:
            gc_log.append((info['generation'], (time.perf_counter() - _t[0]) * 1e3, info['collected']))

    gc.callbacks.append(on_gc)
    if os.environ.get('NPB_PROBE_FREEZE'):
        gc.collect()
        gc.freeze()     # everything alive now (modules, code, tensors) leaves the GC's generations
    n_steps = int(os.environ.get('NPB_PROBE_STEPS', '80'))
    times = []
    for i in range(n_steps):
        t0 = time.perf_counter()
        r = post.postprocess(arm.raw, batch, is_training=False)
        t1 = time.perf_counter()
        helper.validation_step(batch, i, r)
        t2 = time.perf_counter()
        ids = r['panoptic_segmentation_deeplab_ids']
        meta = r['panoptic_segmentation_deeplab_instance_meta']
        t3 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, t3 - t2))
    torch.cuda.synchronize()
    slow = [(i, sum(t) * 1e3) for i, t in enumerate(times) if i > 10 and sum(t) > 3e-3]
    print('slow steps (> 3 ms):', [(i, round(ms, 1)) for i, ms in slow])
    print('gc events > 1 ms (generation, ms, collected):', [(g, round(ms, 1), c) for g, ms, c in gc_log if ms > 1.0])
    print('gc counts per generation:', [sum(1 for g, _, _ in gc_log if g == k) for k in range(3)])
    for i in list(range(0, 16)) + list(range(16, min(n_steps, 80), 8)):
        a, b, c = times[i]
        print(f'step {i:2d}: postprocess {a * 1e6:7.0f} us  validation_step {b * 1e6:7.0f} us  dicts {c * 1e6:6.0f} us')
    steady = sorted(sum(t) for t in times[10:])
    print(f'mean {sum(steady) / len(steady) * 1e3:.3f} ms, median {steady[len(steady) // 2] * 1e3:.3f} ms per step '
          f'({w["B"] / (sum(steady) / len(steady)):.0f} frames/s)')
    print('reserved MB', torch.cuda.memory_reserved() >> 20, 'allocated MB', torch.cuda.memory_allocated() >> 20)


if __name__ == '__main__':
    main()

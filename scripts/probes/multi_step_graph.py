#!/usr/bin/env python
"""How much of the headline step is the gap between two graph replays?  G consecutive steps
captured in ONE graph (the chain of programmatic dependent launches then runs across the step
boundaries), G = 1, 2, 4, 8; CUDA events over the same number of steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from nicr_mt_scene_analysis_b200.graph import CapturedStep  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'nyuv2']
    B = w['B']
    steps = 960
    for G in (1, 2, 4, 8, 1):
        arm = bench.Arm(w, B, dev, 0, fused=True, graph=False, pipeline=True)

        def fn():
            r = None
            for _ in range(G):
                r = arm.eager_step()
            return r

        step = CapturedStep(fn, warmup=3, device=dev).replay
        for _ in range(200 // G):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps // G):
            step()
        arm.pq._flush_deferred()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        res = arm.evaluation.compute(suffix='_deeplab')
        arm.pq.check_status()
        print(f'G={G}: {ms / steps * 1e3:.1f} us per step, {steps * B / ms * 1e3:.0f} frames/s, '
              f"pq {float(res['all_deeplab_pq']):.6f}", flush=True)
        del arm, step
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()

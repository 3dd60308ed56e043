#!/usr/bin/env python
"""Host cost of the drop-in API loop that bench.py reports as `value_api` (synchronous
`postprocess()` + `PanopticTaskHelper.validation_step`, python dicts read every step), with
cProfile: where the host spends its time once the GPU part is ~0.1 ms.

    python scripts/host_profile_api.py [--config nyuv2] [--fused] [--steps 200]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='nyuv2')
    ap.add_argument('--fused', action='store_true')
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--top', type=int, default=28)
    args = ap.parse_args()
    import torch
    import bench
    from nicr_mt_scene_analysis_b200.task_helper import PanopticTaskHelper
    dev = torch.device('cuda', 0)
    w = dict(bench.WORKLOADS[args.config])
    arm = bench.Arm(w, w['B'], dev, 0, fused=False, graph=False)
    post = arm.new_post()
    helper = PanopticTaskHelper(w['C'] + 1, (False,) + arm.is_thing)
    helper.initialize(dev)
    if args.fused:
        post.fuse_evaluation(helper.evaluation)
    batch = dict(arm.batch, panoptic_fullres=arm.tgt_pan, semantic_fullres=arm.tgt_sem)

    def one(i):
        r = post.postprocess(arm.raw, batch, is_training=False)
        helper.validation_step(batch, i, r)
        return r['panoptic_segmentation_deeplab'], r['panoptic_segmentation_deeplab_ids'], \
            r['panoptic_segmentation_deeplab_instance_meta']

    for i in range(5):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        one(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    print(f'{w["name"]} fused={args.fused}: {dt * 1e6:.0f} us per step, {w["B"] / dt:.0f} frames/s')
    prof = cProfile.Profile()
    prof.enable()
    for i in range(args.steps):
        one(i)
    torch.cuda.synchronize()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats('tottime')
    st.print_stats(args.top)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Where does the time of one bench step go?  Times CUDA-graph replays of sub-chains of the
default step in ONE process (same clocks, warm caches, back to back like the real step):
full step, post-processing only, evaluation only, centre detection only, grouping only.
Diagnostic for profiles/README.md; not part of the bench contract.

    python scripts/step_breakdown.py [--config sunrgbd] [--reps 300]
"""
import argparse
import json
import os
import sys
from ctypes import c_float, c_int

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='sunrgbd')
    ap.add_argument('--reps', type=int, default=300)
    args = ap.parse_args()
    import torch
    import bench
    from nicr_mt_scene_analysis_b200 import _lib, testing
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class

    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    w = bench.WORKLOADS[args.config]
    B, C, H, W, K, ORI = w['B'], w['C'], w['H'], w['W'], w['K'], w['ori']
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    pool = 16
    frames = [testing.make_frame(C, H, W, K, seed=1000 + i, with_orientation=ORI, device=dev,
                                 quantize=None) for i in range(pool)]
    data = {k: torch.stack([frames[i % pool][k] for i in range(B)]).contiguous() for k in frames[0]}
    del frames
    batch = testing.make_batch_dict(B, H, W)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance', top_k_instances=w['top_k'])(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori,
        async_results=True)()
    pq = PanopticQuality(C + 1, 0, bench.L, bench.OFFSET, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    ev = PanopticEvaluation(pq, miou)
    inst_out = (data['heat'], data['offset']) + ((data['orientation'],) if ORI else ())
    raw = ((data['logits'], inst_out), (None, None))
    r0 = post.postprocess(raw, batch, is_training=False)
    pan = r0['panoptic_segmentation_deeplab'].clone()
    tabs = r0['_panoptic_instance_tables']
    tgt_pan, tgt_sem = testing.make_eval_targets(pan, bench.L)

    L = _lib.lib()
    lut = _lib.host_lut(is_thing, C)
    sem = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    inst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    hist = torch.empty((B, _lib.MAX_INST, C), dtype=torch.int32, device=dev)
    osum = torch.empty((B, _lib.MAX_INST, 2), dtype=torch.float64, device=dev) if ORI else None
    ws_c = torch.empty(L.npb_instance_centers_workspace_bytes(B, H, W, 3), dtype=torch.uint8, device=dev)
    cyx = torch.empty((B, _lib.MAX_INST, 2), dtype=torch.int32, device=dev)
    ncen = torch.empty((B,), dtype=torch.int32, device=dev)
    cscore = torch.empty((B, _lib.MAX_INST), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)

    def full():
        r = post.postprocess(raw, batch, is_training=False)
        ev.update(r['panoptic_segmentation_deeplab'], tgt_pan, tgt_sem)

    def post_only():
        post.postprocess(raw, batch, is_training=False)

    def eval_only():
        ev.update(pan, tgt_pan, tgt_sem)

    def centers_only():
        _lib.check(L.npb_instance_centers(
            _lib.ptr(data['heat']), c_int(B), c_int(H), c_int(W), c_float(0.1), c_int(3),
            c_int(w['top_k']), None, c_int(0), _lib.ptr(ws_c), _lib.ptr(cyx), _lib.ptr(ncen),
            _lib.ptr(cscore), _lib.ptr(status), _lib.stream_ptr(dev)))

    def group_only():
        _lib.check(L.npb_group_pixels(
            _lib.ptr(data['logits']), None, None, _lib.ptr(data['offset']),
            _lib.ptr(data.get('orientation')), c_int(B), c_int(C), c_int(H), c_int(W), lut,
            tabs.dptr('centers_yx'), tabs.dptr('n_centers'), c_int(1), c_int(0), c_float(0.0),
            _lib.ptr(sem), _lib.ptr(inst), _lib.ptr(hist), _lib.ptr(osum), _lib.stream_ptr(dev)))

    out = {'config': args.config, 'frames': B}
    chains = [('full_step', full), ('post_only', post_only), ('eval_only', eval_only),
              ('centers_only', centers_only), ('group_only', group_only)]
    steps = {name: CapturedStep(fn, warmup=3, device=dev).replay for name, fn in chains}
    for _ in range(200):           # steady-state clocks
        steps['full_step']()
    torch.cuda.synchronize(dev)
    for rnd in range(2):           # second round is the reported one
        for name, _ in chains:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                steps[name]()
            e1.record()
            torch.cuda.synchronize(dev)
            out[name + '_us'] = round(e0.elapsed_time(e1) / args.reps * 1e3, 1)
    out['post_minus_parts_us'] = round(out['post_only_us'] - out['centers_only_us'] - out['group_only_us'], 1)
    out['full_minus_post_eval_us'] = round(out['full_step_us'] - out['post_only_us'] - out['eval_only_us'], 1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()

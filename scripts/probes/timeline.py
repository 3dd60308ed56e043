"""In-situ timeline of one replayed step (needs a library built with -DNPB_TIMELINE, passed through
NPB_LIB_PATH): where do the kernels of the chain start, pass their dependency wait and end?

    python nicr-multitask-scene-analysis_b200/csrc/build.py -DNPB_TIMELINE --out=build/timeline/libnicr_panoptic_b200.so
    NPB_LIB_PATH=build/timeline/libnicr_panoptic_b200.so python scripts/probes/timeline.py --config nyuv2
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from nicr_mt_scene_analysis_b200 import _lib, testing  # noqa: E402
from nicr_mt_scene_analysis_b200.graph import CapturedStep  # noqa: E402
from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,  # noqa: E402
                                                PanopticQualityWithOrientationMAE)
from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class  # noqa: E402

NAMES = {0: 'nms+select', 1: 'group', 2: 'pair(+finalize)', 10: ' p: tables derived', 11: ' p: pixel loop done', 3: 'match', 5: ' m: dense merged', 6: ' m: entries merged',
         7: ' m: segment tables', 8: ' m: matched', 9: ' m: fn/fp + ordered', 4: 'accumulate tail'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='nyuv2')
    ap.add_argument('--frames', type=int, default=0)
    ap.add_argument('--eager', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true')
    args = ap.parse_args()
    w = dict(bench.WORKLOADS[args.config])
    if args.frames:
        w['B'] = args.frames
    dev = torch.device('cuda:0')
    B, C, H, W, K = w['B'], w['C'], w['H'], w['W'], w['K']
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
    frames = [testing.make_frame(C, H, W, K, seed=1000 + i, with_orientation=w['ori'], device=dev,
                                 quantize=None) for i in range(min(B, 16))]
    data = {k: torch.stack([frames[i % len(frames)][k] for i in range(B)]).contiguous() for k in frames[0]}
    batch = testing.make_batch_dict(B, H, W)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance', top_k_instances=w['top_k'])(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori, async_results=True)()
    pq = PanopticQualityWithOrientationMAE(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    ev = PanopticEvaluation(pq, miou)
    inst_out = (data['heat'], data['offset']) + ((data['orientation'],) if w['ori'] else ())
    raw = ((data['logits'], inst_out), (None, None))
    r0 = post.postprocess(raw, batch, is_training=False)
    tgt_pan, tgt_sem = testing.make_eval_targets(r0['panoptic_segmentation_deeplab'], 1 << 16)
    post.fuse_evaluation(ev, pipeline_matching=not args.no_pipeline)
    batch_gt = dict(batch, panoptic_fullres=tgt_pan, semantic_fullres=tgt_sem)

    def eager():
        return post.postprocess(raw, batch_gt, is_training=False)

    step = eager if args.eager else CapturedStep(eager, warmup=3, device=dev).replay
    lib = _lib.lib()
    read = lib.npb_timeline_read
    read.restype = ctypes.c_int
    read.argtypes = [ctypes.c_void_p]
    buf = (ctypes.c_uint64 * (16 * 6))()
    for _ in range(20):
        step()
    read(buf)
    print('build:', lib.npb_build_info().decode(), '| config', w['name'], '| mode', 'eager' if args.eager else 'graph')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        torch.cuda.synchronize()
        step()          # the step before: what the measured one queues behind
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        read(buf)
        v = [buf[i] for i in range(16 * 6)]
        t0 = min(v[6 * k] for k in (0, 1, 2, 3) if v[6 * k] != 2 ** 64 - 1)
        print(f'-- two consecutive steps (times in us since the first CTA start; events: {e0.elapsed_time(e1) * 1e3:.1f} us for the 2nd)')
        for k, name in NAMES.items():
            s_min, s_max, w_min, w_max, e_min, e_max = v[6 * k:6 * k + 6]
            if s_min == 2 ** 64 - 1 and w_min == 2 ** 64 - 1:
                continue
            f = lambda x: f'{(x - t0) / 1e3:8.1f}' if x not in (0, 2 ** 64 - 1) else '       -'
            print(f'{name:18s} start {f(s_min)} .. {f(s_max)}  wait passed {f(w_min)} .. {f(w_max)}  end {f(e_min)} .. {f(e_max)}')


if __name__ == '__main__':
    main()

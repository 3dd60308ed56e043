#!/usr/bin/env python
"""bench.py -- panoptic frames/s (post-processing + merge + mIoU/PQ) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): SUNRGB-D-shaped frames 530x730, 37 classes, batch 64 per
GPU with orientation head, 20 instances per frame, synthetic decoder outputs.
One "step" = one batch through PanopticPostprocessing (centre NMS/top-k, fused arg-max +
offset grouping + votes + orientation, instance table, panoptic ids) followed by the fused
PQ + mIoU update against shifted targets.  Every rank owns its own frames (weak scaling, no
data-path collective); the metric states are all-reduced once, at compute(), inside the
timed region.

Prints ONE JSON line (see the task contract): `value` = frames/s with inputs resident in HBM,
`e2e` = the same through the host-buffer pipeline (pinned host inputs, H2D + D2H inside the
timed region), `roofline` for the dominant kernel (group_pixels_kernel), `cpu_baseline` = the
C oracle port timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs; the default (and the only one the driver runs) is configs[1].
WORKLOADS = {
    'nyuv2': dict(name='nyuv2_480x640_c40_b8', B=8, C=40, H=480, W=640, K=12, top_k=64, ori=False),
    'sunrgbd': dict(name='sunrgbd_530x730_c37_b64_orientation', B=64, C=37, H=530, W=730, K=20,
                    top_k=64, ori=True),
    'scannet': dict(name='scannet_968x1296_c40_b128', B=128, C=40, H=968, W=1296, K=30, top_k=64,
                    ori=False),
    'cityscapes': dict(name='cityscapes_1024x2048_c19_b256_k100', B=256, C=19, H=1024, W=2048,
                       K=100, top_k=100, ori=False),
}
WORKLOAD = dict(WORKLOADS['sunrgbd'])
L = 1 << 16
OFFSET = 256 ** 3
METRIC = 'panoptic frames/s (postproc+merge+mIoU/PQ)'
UNIT = 'frames/s'


def bytes_post_per_frame(C, H, W, orientation=True):
    """compulsory reads of decoder outputs + writes of dense API outputs."""
    return H * W * (4 * C + 4 + 8 + (8 if orientation else 0) + 8 + 1)


def bytes_eval_per_frame(H, W):
    return 17 * H * W


def bytes_group_kernel_per_frame(C, H, W, orientation=True):
    """The share of 8(d) the dominant kernel is responsible for: logits, offsets,
    orientation in; uint8 instance map out."""
    return H * W * (4 * C + 8 + (8 if orientation else 0) + 1)


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.QUERY}',
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append((time.perf_counter(), [x.strip() for x in out.split(',')]))
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken inside [t0, t1] (the timed
        region).  When the timed region is too short to hold 3 samples, the samples of the
        steady-state warm-up that ran the identical load right before it are used as well."""
        inside = [f for t, f in self.samples if t0 is None or t0 <= t <= t1]
        window = 'timed region'
        if len(inside) < 3:
            inside = [f for _, f in self.samples]
            window = 'steady-state warm-up (same load) + timed region'
        sm = sorted(int(s[0]) for s in inside if s and s[0].isdigit())
        mx = max((int(s[1]) for s in inside if len(s) > 1 and s[1].isdigit()), default=None)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in inside for n, v in zip(names, s[2:6])
                          if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons,
                'samples': len(inside), 'window': window}


def oracle_baseline(sample_frames, steps, warmup, threads=None):
    """The CPU path ("port": C restatement of the reference, OpenMP over frames) on a
    bounded sample of the workload: post-processing + PQ + confusion matrix per frame."""
    import numpy as np
    import torch
    import oracle
    from nicr_mt_scene_analysis_b200 import testing
    w = WORKLOAD
    # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on
    oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = oracle.num_threads()
    pool = min(sample_frames, 16)       # distinct frames, cycled (like the GPU arm's pool)
    data = testing.make_batch(pool, w['C'], w['H'], w['W'], w['K'], seed=1,
                              with_orientation=w['ori'], quantize=None)
    idx = np.arange(sample_frames) % pool
    arrs = {k: np.ascontiguousarray(v.numpy()[idx]) for k, v in data.items()}
    is_thing = testing.default_is_thing(w['C'])
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))

    def step():
        r = oracle.panoptic_postprocess(arrs['logits'], arrs['heat'], arrs['offset'],
                                        arrs.get('orientation'), is_thing, has_ori,
                                        top_k=w['top_k'])
        pan = r['panoptic']
        tgt = np.roll(pan, 5, axis=-1)
        sem_t = (tgt // L).astype(np.uint8)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(cores) as ex:       # ctypes releases the GIL
            list(ex.map(lambda b: oracle.pq_compare_and_accumulate(
                pan[b], tgt[b], w['C'] + 1, 0, L, OFFSET, 0), range(sample_frames)))
        oracle.confmat(pan // L, sem_t, w['C'] + 1)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_frames * steps / dt, cores, dt / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = 64
    steps = max(1, min(args.steps, 40))      # bounded: <= 2560 frames of CPU work
    fps, cores, s_per_step = oracle_baseline(sample, steps, min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': s_per_step * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': WORKLOAD['name'], 'sample_frames_per_step': sample},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} frames of the workload per step, C oracle '
                                   '(oracle/panoptic_oracle.c), OpenMP over frames'},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nicr_mt_scene_analysis_b200 import _lib, testing
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQualityWithOrientationMAE)
    from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
    from nicr_mt_scene_analysis_b200.pipeline import PanopticHostPipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference '
                         'for the CPU baseline')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    w = WORKLOAD
    B, C, H, W, K = w['B'], w['C'], w['H'], w['W'], w['K']
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))

    # ---- synthetic decoder outputs, generated on the device (distinct frames per rank) ----
    pool = 16            # distinct frames; the batch cycles them (inputs stay > L2: 4.4 GB)
    ORI = w['ori']
    frames = [testing.make_frame(C, H, W, K, seed=1000 * (rank + 1) + i, with_orientation=ORI,
                                 device=dev, quantize=None) for i in range(pool)]
    data = {k: torch.stack([frames[i % pool][k] for i in range(B)]).contiguous() for k in frames[0]}
    del frames
    batch = testing.make_batch_dict(B, H, W)

    def new_post(**kw):
        return get_postprocessing_class(
            'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
            instance_postprocessing=get_postprocessing_class(
                'instance', top_k_instances=w['top_k'])(),
            semantic_classes_is_thing=is_thing, semantic_class_has_orientation=has_ori, **kw)()

    post = new_post(async_results=True)
    pq = PanopticQualityWithOrientationMAE(C + 1, 0, L, OFFSET, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    evaluation = PanopticEvaluation(pq, miou)

    inst_out = (data['heat'], data['offset']) + ((data['orientation'],) if ORI else ())
    raw = ((data['logits'], inst_out), (None, None))
    # evaluation targets: prediction rolled by 5 px, fixed for the run
    r0 = post.postprocess(raw, batch, is_training=False)
    tgt_pan, tgt_sem = testing.make_eval_targets(r0['panoptic_segmentation_deeplab'], L)
    del r0
    # default: the kernel that writes the panoptic ids also evaluates them (one launch less, the
    # ids are not read back); --no-fuse runs post-processing and evaluation as separate calls
    fused = not args.no_fuse
    if fused:
        post.fuse_evaluation(evaluation)
    batch_gt = dict(batch, panoptic_fullres=tgt_pan, semantic_fullres=tgt_sem) if fused else batch
    # nms, select, group, finalize, write (+) pair_count, match, accumulate
    KERNELS_PER_STEP = 7 if fused else 8

    def eager_step():
        r = post.postprocess(raw, batch_gt, is_training=False)
        if not r.get('_panoptic_evaluation_fused'):
            evaluation.update(r['panoptic_segmentation_deeplab'], tgt_pan, tgt_sem)
        return r

    launch_mode = 'eager'
    step = eager_step
    if not args.no_graph:
        # the step is 7-8 short kernels: capture them once, replay with one launch per step
        from nicr_mt_scene_analysis_b200.graph import CapturedStep
        try:
            step = CapturedStep(eager_step, warmup=3, device=dev).replay
            launch_mode = 'cuda graph replay'
        except Exception as exc:       # keep the benchmark alive, say what happened
            print(f'[bench] CUDA graph capture failed ({exc!r}); issuing steps eagerly',
                  file=sys.stderr)
            torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # W warm-up steps (at least 3), extended until the GPU has been busy for ~1 s so that the
    # timed region sees steady-state clocks even when it is only a few milliseconds long
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    t_warm = time.perf_counter()
    n_warm = 0
    while n_warm < max(args.warmup, 3) or time.perf_counter() - t_warm < 1.0:
        step()
        n_warm += 1
        if n_warm % 16 == 0:
            torch.cuda.synchronize(dev)
    evaluation.compute(suffix='_deeplab')   # warm-up of the metric all-reduce (NCCL connections)
    evaluation.reset()
    barrier()

    # ---- timed region: device-resident inputs ------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    region0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        last = step()
    results = evaluation.compute(suffix='_deeplab')        # one all-reduce of the states
    e1.record()
    barrier()
    region1 = time.perf_counter()
    clocks.__exit__(None, None, None)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    frames_total = B * args.steps * world
    value = frames_total / (ms * 1e-3)
    last['_panoptic_instance_tables'].wait()        # per-frame status words of the last step
    pq.check_status()

    # ---- dominant kernel in isolation: npb_group_pixels (CUDA events on its stream) --------------
    from ctypes import c_float, c_int
    tabs = last['_panoptic_instance_tables']
    sem = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    inst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    hist = torch.empty((B, _lib.MAX_INST, C), dtype=torch.int32, device=dev)
    osum = torch.empty((B, _lib.MAX_INST, 2), dtype=torch.float64, device=dev) if ORI else None
    lut = _lib.host_lut(is_thing, C)

    def group_only():
        _lib.check(_lib.lib().npb_group_pixels(
            _lib.ptr(data['logits']), None, None, _lib.ptr(data['offset']),
            _lib.ptr(data.get('orientation')), c_int(B), c_int(C), c_int(H), c_int(W), lut,
            tabs.dptr('centers_yx'), tabs.dptr('n_centers'), c_int(1), c_int(0), c_float(0.0),
            _lib.ptr(sem), _lib.ptr(inst), _lib.ptr(hist), _lib.ptr(osum), _lib.stream_ptr(dev)))

    for _ in range(3):
        group_only()
    reps = max(args.steps, 5)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    k0.record()
    for _ in range(reps):
        group_only()
    k1.record()
    torch.cuda.synchronize(dev)
    kernel_ms = k0.elapsed_time(k1) / reps      # includes the two small memsets of the call
    peak, peak_src = measured_peak_gbs()
    kbytes = bytes_group_kernel_per_frame(C, H, W, ORI) * B
    achieved = kbytes / (kernel_ms * 1e-3) / 1e9

    # ---- end to end: pinned host buffers in, panoptic ids in host memory out -------------------
    e2e = None
    if not args.no_e2e:
        host_in = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v)
                   for k, v in data.items()}
        host_tgt = {'panoptic': torch.empty(tgt_pan.shape, dtype=torch.int64, pin_memory=True).copy_(tgt_pan),
                    'semantic': torch.empty(tgt_sem.shape, dtype=torch.uint8, pin_memory=True).copy_(tgt_sem)}
        # two sets of host result buffers: batch k+1 is enqueued before the python structures of
        # batch k are built, so the host->device link never waits for the host
        outs = [{'panoptic_segmentation_deeplab': torch.empty((B, H, W), dtype=torch.int64, pin_memory=True),
                 'panoptic_segmentation_deeplab_instance_idx': torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)}
                for _ in range(2)]
        evaluation.reset()
        pipe = PanopticHostPipeline(new_post(async_results=True), evaluation, chunk_frames=8, device=dev)
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_steps_run(n):
            pending = None
            for i in range(n):
                o = pipe.run(host_in, batch, host_tgt, out=dict(outs[i % 2]))
                if pending is not None:
                    PanopticHostPipeline.finish(pending, with_orientation=ORI)   # blocks on batch i-1 only
                pending = o
            return PanopticHostPipeline.finish(pending, with_orientation=ORI)

        e2e_steps_run(2)
        evaluation.reset()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        t0.record()
        e2e_steps_run(e2e_steps)
        evaluation.compute(suffix='_deeplab')
        t1.record()
        barrier()
        wall = time.perf_counter() - wall0
        ems = torch.tensor([max(t0.elapsed_time(t1), wall * 1e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {'value': B * e2e_steps * world / (float(ems.item()) * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': pipe.h2d_bytes, 'd2h_bytes_per_step': pipe.d2h_bytes,
               'steps': e2e_steps, 'chunk_frames': 8}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, cores, _ = oracle_baseline(64, 10, 1)
        cpu = {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '64 frames of the workload x 10 steps (640 frames), C oracle '
                         '(oracle/panoptic_oracle.c), OpenMP over frames'}

    if rank == 0:
        bpf = bytes_post_per_frame(C, H, W, ORI) + bytes_eval_per_frame(H, W)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'warmup_steps_run': n_warm,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'frames_per_gpu_per_step': B, 'classes': C,
                       'height': H, 'width': W, 'instances_per_frame': K,
                       'parallelism': f'frames sharded over {world} GPU(s), metric states '
                                      'all-reduced at compute()',
                       'l2_policy': f'inputs ({B * bytes_post_per_frame(C, H, W, ORI) / 1e9:.1f} GB per step) '
                                    'larger than L2, no flush needed',
                       'launch': launch_mode,
                       'evaluation': 'fused into the kernel that writes the panoptic ids' if fused
                       else 'separate call on the written ids'},
            'clocks': clocks.summary(region0, region1),
            'e2e': e2e,
            'gpu_launches': KERNELS_PER_STEP * args.steps,
            'roofline': {'bound': 'hbm', 'kernel': 'group_pixels_kernel<4,logits,%s>' % ('orientation' if ORI else 'no orientation'),
                         'achieved': achieved, 'peak': peak, 'peak_source': peak_src,
                         'unit': 'GB/s', 'frac': achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                         # capture of this command (profiles/r01_group_pixels_ncu_raw.csv)
                         'traffic': 4.0221e9 if args.config == 'sunrgbd' and not args.frames else None,
                         'kernel_ms': kernel_ms, 'algorithmic_bytes_per_launch': kbytes},
            'roofline_path': {'bytes_per_frame': bpf,
                              'achieved': value / world * bpf / 1e9, 'unit': 'GB/s',
                              'frac': value / world * bpf / 1e9 / peak},
            'cpu_baseline': cpu,
            'quality': {'all_pq': float(results['all_deeplab_pq']),
                        'miou': float(results['semantic_deeplab_miou'])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='sunrgbd', choices=sorted(WORKLOADS),
                    help='BASELINE.json shape (default: configs[1], the metric\'s configuration)')
    ap.add_argument('--frames', type=int, default=0, help='override frames per GPU per step')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='issue every step from Python')
    ap.add_argument('--no-fuse', action='store_true',
                    help='post-processing and evaluation as separate calls (8 kernels per step)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.config])
    if args.frames:
        WORKLOAD['B'] = args.frames
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

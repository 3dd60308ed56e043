#!/usr/bin/env python
"""bench.py -- panoptic frames/s (post-processing + merge + mIoU/PQ) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload = the configuration BASELINE.json's metric is quoted on: NYUv2-shaped frames
480x640, 40 classes, batch 8 per GPU, 12 instances per frame, synthetic decoder outputs
(SURVEY.md 8d).  One "step" = one batch through PanopticPostprocessing (centre NMS/top-k, fused
arg-max + offset grouping + votes, instance table, panoptic ids) with the PQ + mIoU update
against shifted targets fused into the id writer.  Every rank owns its own frames (weak
scaling, no data-path collective).  The timed region is the K steps including ALL their GPU work
(the matcher of the last batch, which the pipelined evaluation leaves pending, is flushed inside
it); compute() -- once per validation epoch: one all-reduce of the metric states per dtype + host
arithmetic -- follows and is reported as `epoch_end` (the 50 k-frame evaluation of extra.configs
keeps it inside its timed region, as BASELINE.json configs[4] asks).

Prints ONE JSON line (task contract):
  value        frames/s, inputs resident in HBM, CUDA-graph replay of the step
  e2e          the same through the host-buffer pipeline (pinned host inputs, H2D + D2H timed)
  value_api    eager, synchronous `postprocess()` + `PanopticTaskHelper.validation_step` with the
               python dicts of the reference API read every step (what a drop-in caller gets)
  roofline     dominant kernel (group_pixels_kernel) against the measured HBM peak
  cpu_baseline the UNMODIFIED reference (baseline/_ref, torch CPU path) on this box's host cores
               on a bounded sample (`kind: "reference"`); the C oracle port as a second key
  extra.configs  the other BASELINE.json shapes (SUNRGB-D, ScanNet, Cityscapes, 50 k-frame
               evaluation); at N > 1 their batches are SPLIT over the ranks (strong scaling)
`--impl reference` times the reference alone (rank 0), same metric / config keys.
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

# BASELINE.json configs.  B = frames per GPU per step of the headline run; `total` = the batch
# BASELINE.json names for the sharded shapes (split over the ranks in extra.configs at N > 1).
WORKLOADS = {
    'nyuv2': dict(name='nyuv2_480x640_c40_b8', B=8, C=40, H=480, W=640, K=12, top_k=64, ori=False),
    'sunrgbd': dict(name='sunrgbd_530x730_c37_b64_orientation', B=64, C=37, H=530, W=730, K=20,
                    top_k=64, ori=True),
    'scannet': dict(name='scannet_968x1296_c40_b128', B=128, C=40, H=968, W=1296, K=30, top_k=64,
                    ori=False),
    'cityscapes': dict(name='cityscapes_1024x2048_c19_b256_k100', B=256, C=19, H=1024, W=2048,
                       K=100, top_k=100, ori=False),
}
DEFAULT_CONFIG = 'nyuv2'        # "@480x640": the configuration the metric is quoted on
WORKLOAD = dict(WORKLOADS[DEFAULT_CONFIG])
EVAL_FRAMES = 50000             # BASELINE.json configs[4]
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


def ncu_traffic(config, frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from
    the committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written
    by scripts/ncu_traffic.py from the raw csv pages) -- None when no capture of this exact
    workload is committed."""
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        e = table.get(f'{config}_b{frames}')
        return (float(e['dram_bytes']), e['source']) if e else (None, None)
    except Exception:
        return None, None


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


# ------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref) and the C oracle port
# ------------------------------------------------------------------------------------------------
def oracle_baseline(sample_frames, steps, warmup, threads=None):
    """The C restatement ("port", OpenMP over frames) on a bounded sample of the workload:
    post-processing + PQ + confusion matrix per frame."""
    import numpy as np
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


def reference_sample_frames(w):
    """Frames per reference step: a whole batch of the workload where that stays within a few
    seconds of CPU work per step, else a bounded sample of it (0.5 s per
    NYUv2 frame, 1 s per SUNRGB-D frame, 5 s per ScanNet frame, 30 s per Cityscapes frame on 8 cores)."""
    return {'nyuv2': 8, 'sunrgbd': 8, 'scannet': 2, 'cityscapes': 1}.get(
        w['name'].split('_')[0], min(w['B'], 8))


def run_reference(args):
    """`--impl reference`: the unmodified reference's own CPU implementation of the path
    (PanopticPostprocessing.postprocess + PanopticQuality.update + mIoU.update from
    baseline/_ref) on the host cores; the C oracle port only if the reference is not installed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import reference_arm
    w = WORKLOAD
    sample = args.frames or reference_sample_frames(w)
    warm = max(1, min(args.warmup, 1))
    if reference_arm.available() and not args.port:
        steps = max(1, min(args.steps, args.reference_max_steps))
        r = reference_arm.time_reference(w, sample, steps, warm)
        fps, cores, s_per_step = r['value'], r['cores'], r['s_per_step']
        kind = 'reference'
        what = (f'{sample} frames of the workload per step x {steps} steps, UNMODIFIED reference '
                f'(baseline/_ref: PanopticPostprocessing.postprocess + PanopticQuality.update, '
                f'its own pool of {min(cores, 32)} single-thread workers, + mIoU.update), torch CPU, '
                f'{cores} intra-op threads')
        quality = {'all_pq': r['all_pq'], 'miou': r['miou']}
    else:
        steps = max(1, min(args.steps, 40))
        sample = args.frames or 64
        fps, cores, s_per_step = oracle_baseline(sample, steps, warm)
        kind = 'port'
        what = (f'{sample} frames of the workload per step x {steps} steps, C oracle '
                '(oracle/panoptic_oracle.c), OpenMP over frames')
        quality = None
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': s_per_step * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': w['name'], 'frames_per_gpu_per_step': w['B'], 'classes': w['C'],
                   'height': w['H'], 'width': w['W'], 'instances_per_frame': w['K'],
                   'sample_frames_per_step': sample},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': what},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'quality': quality,
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(config, steps, port=False, timeout=600):
    """cpu_baseline of the GPU arm: `bench.py --impl reference` in its own process (its thread
    settings and its worker pool stay out of this one) -> its parsed line, or None."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', config,
           '--steps', str(steps), '--warmup', '1'] + (['--port'] if port else [])
    env = {k: v for k, v in os.environ.items()
           if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS')}
    env['CUDA_VISIBLE_DEVICES'] = ''
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        print('[bench] reference arm printed no line:', out.stderr[-400:], file=sys.stderr)
    except Exception as exc:
        print(f'[bench] reference arm failed: {exc!r}', file=sys.stderr)
    return None


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Arm:
    """Post-processing + evaluation of one workload on one device: synthetic decoder outputs
    resident in HBM, the step (eager or captured), timing helpers."""

    def __init__(self, w, B, dev, rank, fused=True, graph=True, pipeline=True):
        import torch
        from nicr_mt_scene_analysis_b200 import testing
        from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion,
                                                        PanopticEvaluation,
                                                        PanopticQualityWithOrientationMAE)
        from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
        self.torch = torch
        self.w, self.B, self.dev = w, B, dev
        C, H, W, K = w['C'], w['H'], w['W'], w['K']
        self.is_thing = testing.default_is_thing(C)
        self.has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(self.is_thing))
        ORI = w['ori']
        # distinct frames per rank; the batch cycles a pool of 16 (inputs stay larger than L2)
        pool = min(16, B)
        frames = [testing.make_frame(C, H, W, K, seed=1000 * (rank + 1) + i, with_orientation=ORI,
                                     device=dev, quantize=None) for i in range(pool)]
        self.data = {k: torch.stack([frames[i % pool][k] for i in range(B)]).contiguous()
                     for k in frames[0]}
        del frames
        self.batch = testing.make_batch_dict(B, H, W)

        def new_post(**kw):
            return get_postprocessing_class(
                'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
                instance_postprocessing=get_postprocessing_class(
                    'instance', top_k_instances=w['top_k'])(),
                semantic_classes_is_thing=self.is_thing,
                semantic_class_has_orientation=self.has_ori, **kw)()

        self.new_post = new_post
        self.post = new_post(async_results=True)
        self.pq = PanopticQualityWithOrientationMAE(C + 1, 0, L, OFFSET, (False,) + self.is_thing,
                                                    device=dev)
        self.miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
        self.evaluation = PanopticEvaluation(self.pq, self.miou)
        d = self.data
        inst_out = (d['heat'], d['offset']) + ((d['orientation'],) if ORI else ())
        self.raw = ((d['logits'], inst_out), (None, None))
        # evaluation targets: prediction rolled by 5 px, fixed for the run
        r0 = self.post.postprocess(self.raw, self.batch, is_training=False)
        self.tgt_pan, self.tgt_sem = testing.make_eval_targets(r0['panoptic_segmentation_deeplab'], L)
        del r0
        self.fused = fused
        self.pipelined = bool(fused and pipeline)
        if fused:
            # validation loop: the PQ matcher of a batch runs next to the centre detection +
            # grouping of the next one (and inside compute() for the last batch)
            self.post.fuse_evaluation(self.evaluation, pipeline_matching=self.pipelined)
        self.batch_gt = dict(self.batch, panoptic_fullres=self.tgt_pan,
                             semantic_fullres=self.tgt_sem) if fused else self.batch
        # fused: centre NMS + selection | arg-max + grouping | instance tables + ids + pixel pass |
        # matcher + frame accumulation; separate calls: + finalize, id writer
        self.kernels_per_step = 4 if fused else 6
        self.launch_mode = 'eager'
        self.step = self.eager_step
        if graph:
            from nicr_mt_scene_analysis_b200.graph import CapturedStep
            try:
                self.step = CapturedStep(self.eager_step, warmup=3, device=dev).replay
                self.launch_mode = 'cuda graph replay'
            except Exception as exc:       # keep the benchmark alive, say what happened
                print(f'[bench] CUDA graph capture failed ({exc!r}); issuing steps eagerly',
                      file=sys.stderr)
                torch.cuda.synchronize(dev)

    def eager_step(self):
        r = self.post.postprocess(self.raw, self.batch_gt, is_training=False)
        if not r.get('_panoptic_evaluation_fused'):
            self.evaluation.update(r['panoptic_segmentation_deeplab'], self.tgt_pan, self.tgt_sem)
        return r

    def warm(self, min_steps, min_seconds):
        """At least `min_steps` (>= 3) untimed steps, extended until the GPU has been busy for
        `min_seconds` so that a short timed region sees steady-state clocks."""
        t0 = time.perf_counter()
        n = 0
        while n < max(min_steps, 3) or time.perf_counter() - t0 < min_seconds:
            self.step()
            n += 1
            if n % 16 == 0:
                self.torch.cuda.synchronize(self.dev)
        self.evaluation.compute(suffix='_deeplab')   # also warms the metric all-reduce (NCCL)
        self.evaluation.reset()
        return n

    def time_steps(self, steps, barrier, allreduce_max):
        """-> (ms of `steps` steps, max over ranks; results; last result dict; host perf_counter
        interval of the region; ms of compute()).  The timed region ends when ALL the GPU work of
        the `steps` batches is done -- including the matcher of the last batch, which the pipelined
        evaluation leaves pending.  compute() (once per validation epoch: the all-reduce of the
        metric states + host arithmetic) follows and is reported separately."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            last = self.step()
        self.pq._flush_deferred()
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = allreduce_max(e0.elapsed_time(e1))
        c0 = time.perf_counter()
        results = self.evaluation.compute(suffix='_deeplab')        # one all-reduce of the states
        torch.cuda.synchronize(self.dev)
        compute_ms = allreduce_max((time.perf_counter() - c0) * 1e3)
        last['_panoptic_instance_tables'].wait()        # per-frame status words of the last step
        self.pq.check_status()
        return ms, results, last, (t0, t1), compute_ms

    def kernel_only(self, tabs, reps):
        """The dominant kernel in isolation: npb_group_pixels, CUDA events on its stream."""
        torch = self.torch
        from ctypes import c_float, c_int
        from nicr_mt_scene_analysis_b200 import _lib
        w, B, dev, d = self.w, self.B, self.dev, self.data
        C, H, W = w['C'], w['H'], w['W']
        sem = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        inst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        hist = torch.empty((B, _lib.MAX_INST, C), dtype=torch.int32, device=dev)
        osum = torch.empty((B, _lib.MAX_INST, 2), dtype=torch.float64, device=dev) if w['ori'] else None
        lut = _lib.host_lut(self.is_thing, C)

        def group_only():
            _lib.check(_lib.lib().npb_group_pixels(
                _lib.ptr(d['logits']), None, None, _lib.ptr(d['offset']),
                _lib.ptr(d.get('orientation')), c_int(B), c_int(C), c_int(H), c_int(W), lut,
                tabs.dptr('centers_yx'), tabs.dptr('n_centers'), c_int(1), c_int(0), c_float(0.0),
                _lib.ptr(sem), _lib.ptr(inst), _lib.ptr(hist), _lib.ptr(osum), _lib.stream_ptr(dev)))

        for _ in range(3):
            group_only()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        k0.record()
        for _ in range(reps):
            group_only()
        k1.record()
        torch.cuda.synchronize(dev)
        return k0.elapsed_time(k1) / reps      # includes the two small memsets of the call

    def path_bytes_per_frame(self):
        w = self.w
        return bytes_post_per_frame(w['C'], w['H'], w['W'], w['ori']) + bytes_eval_per_frame(w['H'], w['W'])


def measure_extra_config(name, world, rank, dev, barrier, allreduce_max, peak):
    """One of the other BASELINE.json shapes, briefly: the named batch SPLIT over the ranks
    (strong scaling at N > 1), graph replay, inputs resident."""
    import torch
    w = dict(WORKLOADS[name])
    total = w['B']
    B = max(1, total // world)
    arm = Arm(w, B, dev, rank, fused=True, graph=True)
    arm.warm(3, 0.25)
    step_s = 1e-3 * B * arm.path_bytes_per_frame() / (peak * 1e6 * 0.5)     # rough: half of peak
    steps = int(max(5, min(200, 0.25 / max(step_s, 1e-6))))
    ms, results, last, _, _ = arm.time_steps(steps, barrier, allreduce_max)
    kernel_ms = arm.kernel_only(last['_panoptic_instance_tables'], max(5, min(steps, 50)))
    fps = B * world * steps / (ms * 1e-3)
    bpf = arm.path_bytes_per_frame()
    kbytes = bytes_group_kernel_per_frame(w['C'], w['H'], w['W'], w['ori']) * B
    out = {'workload': w['name'], 'batch_total': B * world, 'frames_per_gpu_per_step': B,
           'scaling': 'strong (the named batch split over the ranks)' if world > 1 else 'single GPU',
           'steps': steps, 'ms_per_step': ms / steps, 'value': fps, 'unit': UNIT,
           'roofline_path_frac': fps / world * bpf / 1e9 / peak,
           'dominant_kernel_frac': kbytes / (kernel_ms * 1e-3) / 1e9 / peak,
           'dominant_kernel_ms': kernel_ms,
           'quality': {'all_pq': float(results['all_deeplab_pq']),
                       'miou': float(results['semantic_deeplab_miou'])}}
    if name == 'sunrgbd':
        # the drop-in validation loop on the 64-frame batch round 1 was judged on (eager,
        # synchronous postprocess() + PanopticTaskHelper.validation_step, ids / meta dicts read
        # every step, orientation MAAE included): this rank's frames per second
        api = measure_value_api(arm, 50)
        out['value_api'] = {k: dict(v, frames_per_step=B) for k, v in api.items()}
    del arm
    torch.cuda.empty_cache()
    return out


def measure_eval_50k(world, rank, dev, barrier, allreduce_max, peak):
    """BASELINE.json configs[4]: PQ + mIoU accumulation over 50 k synthetic 480x640 frames (split
    over the ranks), metric states all-reduced at compute() inside the timed region."""
    import torch
    from nicr_mt_scene_analysis_b200 import testing
    from nicr_mt_scene_analysis_b200.graph import CapturedStep
    from nicr_mt_scene_analysis_b200.metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                                                    PanopticQuality)
    w = WORKLOADS['nyuv2']
    C, H, W, K = w['C'], w['H'], w['W'], w['K']
    B = 256
    arm = Arm(w, 32, dev, rank, fused=False, graph=False)       # 32 distinct predicted frames
    preds = arm.eager_step()['panoptic_segmentation_deeplab'].clone()
    del arm
    pred = preds[torch.arange(B, device=dev) % preds.shape[0]].contiguous()
    tgt, tgt_sem = testing.make_eval_targets(pred, L)
    is_thing = testing.default_is_thing(C)
    pq = PanopticQuality(C + 1, 0, L, OFFSET, (False,) + is_thing, device=dev)
    miou = MeanIntersectionOverUnion(C + 1, ignore_first_class=True, device=dev)
    ev = PanopticEvaluation(pq, miou)
    steps = -(-EVAL_FRAMES // (B * world))
    step = CapturedStep(lambda: ev.update(pred, tgt, tgt_sem), warmup=3, device=dev).replay
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.25:
        step()
    ev.compute()
    ev.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    res = ev.compute()
    e1.record()
    barrier()
    ms = allreduce_max(e0.elapsed_time(e1))
    pq.check_status()
    frames = steps * B * world
    fps = frames / (ms * 1e-3)
    out = {'workload': f'eval_480x640_c40_{EVAL_FRAMES}_frames', 'frames': frames,
           'frames_per_gpu_per_step': B, 'steps': steps, 'ms_total': ms, 'value': fps, 'unit': UNIT,
           'scaling': 'strong (frames split over the ranks, one all-reduce at compute())'
           if world > 1 else 'single GPU',
           'roofline_path_frac': fps / world * 17 * H * W / 1e9 / peak,
           'quality': {'all_pq': float(res['all_pq']), 'miou': float(res['semantic_miou'])}}
    del pred, tgt, tgt_sem, preds
    torch.cuda.empty_cache()
    return out


def measure_value_api(arm, steps):
    """What a drop-in caller gets: eager, synchronous `postprocess()` (python dicts built every
    step) + `PanopticTaskHelper.validation_step`, written like the reference's validation loop
    (task_helper/panoptic.py:87-182); inputs and targets resident on the device.  Reported twice:
    the strict reference call structure, and with the one extra line
    `post.fuse_evaluation(helper.evaluation)`."""
    import torch
    from nicr_mt_scene_analysis_b200.task_helper import PanopticTaskHelper
    w, dev = arm.w, arm.dev
    out = {}
    for mode in ('drop_in', 'fuse_evaluation'):
        post = arm.new_post()
        helper = PanopticTaskHelper(w['C'] + 1, (False,) + arm.is_thing)
        helper.initialize(dev)
        if mode == 'fuse_evaluation':
            post.fuse_evaluation(helper.evaluation)
        batch = dict(arm.batch, panoptic_fullres=arm.tgt_pan, semantic_fullres=arm.tgt_sem)

        def one(i):
            r = post.postprocess(arm.raw, batch, is_training=False)
            helper.validation_step(batch, i, r)
            # the entries the reference's callers read after a validation step
            ids = r['panoptic_segmentation_deeplab_ids']
            meta = r['panoptic_segmentation_deeplab_instance_meta']
            return r['panoptic_segmentation_deeplab'], ids, meta

        # steady state: the follow-up of the metric keeps the tensors of a few updates alive and
        # recycles pinned status / table buffers -- the caching allocator and those pools have
        # grown to their final size after ~2 x FOLLOW_UP_DEPTH + 2 steps
        for i in range(10):
            one(i)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        per_step = []
        for i in range(steps):
            s0 = time.perf_counter()
            one(i)
            per_step.append(time.perf_counter() - s0)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        helper.validation_epoch_end()
        per_step.sort()
        out[mode] = {'value': arm.B * steps / dt, 'unit': UNIT, 'ms_per_step': dt / steps * 1e3,
                     'ms_per_step_median': per_step[len(per_step) // 2] * 1e3,
                     'ms_per_step_max': per_step[-1] * 1e3, 'steps': steps}
    return out


def pcie_ceiling(dev, barrier, nbytes=1 << 30, reps=4):
    """Pinned host -> device copy bandwidth of THIS rank while every rank copies at the same
    time: the ceiling of `e2e` (GB/s)."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def bind_to_gpu_cores(local_rank):
    """Run this rank (and allocate its pinned buffers) on the cores next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * i + b for i, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nicr_mt_scene_analysis_b200.pipeline import PanopticHostPipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference '
                         'for the CPU baseline')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    bound_cores = bind_to_gpu_cores(local_rank) if world > 1 and not os.environ.get('NPB_BENCH_NO_BIND') else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allreduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = WORKLOAD
    B, C, H, W, K = w['B'], w['C'], w['H'], w['W'], w['K']
    ORI = w['ori']
    peak, peak_src = measured_peak_gbs()
    arm = Arm(w, B, dev, rank, fused=not args.no_fuse, graph=not args.no_graph,
              pipeline=not args.no_pipeline)

    # ---- timed region: device-resident inputs ------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    n_warm = arm.warm(args.warmup, 1.0)
    barrier()
    ms, results, last, (region0, region1), compute_ms = arm.time_steps(args.steps, barrier,
                                                                       allreduce_max)
    clocks.__exit__(None, None, None)
    value = B * args.steps * world / (ms * 1e-3)

    kernel_ms = arm.kernel_only(last['_panoptic_instance_tables'], max(min(args.steps, 200), 5))
    kbytes = bytes_group_kernel_per_frame(C, H, W, ORI) * B
    achieved = kbytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.config, B)

    # ---- end to end: pinned host buffers in, panoptic ids in host memory out -------------------
    e2e = None
    if not args.no_e2e:
        data, tgt_pan, tgt_sem = arm.data, arm.tgt_pan, arm.tgt_sem
        host_in = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v)
                   for k, v in data.items()}
        host_tgt = {'panoptic': torch.empty(tgt_pan.shape, dtype=torch.int64, pin_memory=True).copy_(tgt_pan),
                    'semantic': torch.empty(tgt_sem.shape, dtype=torch.uint8, pin_memory=True).copy_(tgt_sem)}
        # two sets of host result buffers: batch k+1 is enqueued before the python structures of
        # batch k are built, so the host->device link never waits for the host
        outs = [{'panoptic_segmentation_deeplab': torch.empty((B, H, W), dtype=torch.int64, pin_memory=True),
                 'panoptic_segmentation_deeplab_instance_idx': torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)}
                for _ in range(2)]
        arm.evaluation.reset()
        chunk = min(8, B)
        pipe = PanopticHostPipeline(arm.new_post(async_results=True), arm.evaluation,
                                    chunk_frames=chunk, device=dev)
        # enough batches for ~0.3 s of PCIe time, at least 5
        h2d_est = sum(v.numel() * v.element_size() for v in host_in.values())
        e2e_steps = max(5, min(args.steps, int(0.3 * 50e9 / max(h2d_est, 1))))

        def e2e_steps_run(n):
            pending = None
            for i in range(n):
                o = pipe.run(host_in, arm.batch, host_tgt, out=dict(outs[i % 2]))
                if pending is not None:
                    PanopticHostPipeline.finish(pending, with_orientation=ORI)   # blocks on batch i-1 only
                pending = o
            return PanopticHostPipeline.finish(pending, with_orientation=ORI)

        e2e_steps_run(2)
        arm.evaluation.reset()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        t0.record()
        e2e_steps_run(e2e_steps)
        arm.evaluation.compute(suffix='_deeplab')
        t1.record()
        barrier()
        wall = time.perf_counter() - wall0
        ems = allreduce_max(max(t0.elapsed_time(t1), wall * 1e3))
        ceiling = pcie_ceiling(dev, barrier)
        ceil_min = -allreduce_max(-ceiling)
        e2e = {'value': B * e2e_steps * world / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': pipe.h2d_bytes, 'd2h_bytes_per_step': pipe.d2h_bytes,
               'steps': e2e_steps, 'chunk_frames': chunk,
               # what the host link gave this job and what it can give with all ranks copying
               'h2d_gbs_per_rank': pipe.h2d_bytes * e2e_steps / (ems * 1e-3) / 1e9,
               'h2d_ceiling_gbs_per_rank': ceil_min,
               'h2d_ceiling': f'1 GiB pinned->device copies, all {world} rank(s) at the same time, '
                              'slowest rank',
               'rank_bound_to_gpu_cores': bound_cores}
        del host_in, host_tgt, outs, pipe

    value_api = None
    if not args.no_api:
        value_api = measure_value_api(arm, max(5, min(args.steps, 50)))
        if world > 1:       # whole-job figure: every rank runs its own loop
            for m in value_api.values():
                m['value'] = -allreduce_max(-m['value']) * world

    bpf = arm.path_bytes_per_frame()
    quality = {'all_pq': float(results['all_deeplab_pq']),
               'miou': float(results['semantic_deeplab_miou'])}
    clock_summary = clocks.summary(region0, region1)
    launch_mode, kernels_per_step, fused = arm.launch_mode, arm.kernels_per_step, arm.fused
    pipelined = arm.pipelined
    del arm, last
    torch.cuda.empty_cache()

    extra = None
    if not args.no_extra:
        extra = {'configs': {}}
        for name in ('sunrgbd', 'scannet', 'cityscapes'):
            if name == args.config:
                continue
            try:
                extra['configs'][name] = measure_extra_config(name, world, rank, dev, barrier,
                                                              allreduce_max, peak)
            except Exception as exc:
                extra['configs'][name] = {'error': repr(exc)}
                torch.cuda.empty_cache()
        try:
            extra['configs']['eval50k'] = measure_eval_50k(world, rank, dev, barrier,
                                                           allreduce_max, peak)
        except Exception as exc:
            extra['configs']['eval50k'] = {'error': repr(exc)}

    cpu = cpu_port = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = reference_subprocess(args.config, 3)
        if ref is not None:
            cpu = ref['cpu_baseline']
            cpu['quality'] = ref.get('quality')
        port = reference_subprocess(args.config, 10, port=True)
        if port is not None:
            cpu_port = port['cpu_baseline']
            if cpu is None:
                cpu = cpu_port

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'warmup_steps_run': n_warm,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            # once per validation epoch, after the timed steps: all-reduce of the metric states
            # (one per dtype) + the host arithmetic of PanopticQuality / mIoU .compute()
            'epoch_end': {'compute_ms': compute_ms,
                          'value_including_compute': B * args.steps * world / ((ms + compute_ms) * 1e-3)},
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'frames_per_gpu_per_step': B, 'classes': C,
                       'height': H, 'width': W, 'instances_per_frame': K,
                       'parallelism': f'frames sharded over {world} GPU(s), metric states '
                                      'all-reduced at compute()',
                       'l2_policy': f'inputs ({B * bytes_post_per_frame(C, H, W, ORI) / 1e9:.2f} GB per step) '
                                    'larger than L2 (126 MB), no flush needed',
                       'launch': launch_mode,
                       'evaluation': ('fused into the kernel that writes the panoptic ids' +
                                      ('; PQ matcher of batch k overlapped with centre detection + '
                                       'grouping of batch k+1 (last batch: inside compute())'
                                       if pipelined else '')) if fused
                       else 'separate call on the written ids'},
            'clocks': clock_summary,
            'e2e': e2e,
            'value_api': value_api,
            'gpu_launches': kernels_per_step * args.steps,
            'roofline': {'bound': 'hbm', 'kernel': 'group_pixels_kernel<4,logits,%s>' % ('orientation' if ORI else 'no orientation'),
                         'achieved': achieved, 'peak': peak, 'peak_source': peak_src,
                         'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'kernel_ms': kernel_ms, 'algorithmic_bytes_per_launch': kbytes},
            'roofline_path': {'bytes_per_frame': bpf,
                              'achieved': value / world * bpf / 1e9, 'unit': 'GB/s',
                              'frac': value / world * bpf / 1e9 / peak},
            'cpu_baseline': cpu,
            'cpu_baseline_port': cpu_port,
            'quality': quality,
            'extra': extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default=DEFAULT_CONFIG, choices=sorted(WORKLOADS),
                    help='BASELINE.json shape; default nyuv2 = 480x640, 40 classes, 8 frames per '
                         'GPU per step, the configuration the metric is quoted on')
    ap.add_argument('--frames', type=int, default=0, help='override frames per GPU per step')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-api', action='store_true', help='skip the value_api measurement')
    ap.add_argument('--no-extra', action='store_true', help='skip extra.configs')
    ap.add_argument('--no-graph', action='store_true', help='issue every step from Python')
    ap.add_argument('--no-fuse', action='store_true',
                    help='post-processing and evaluation as separate calls')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='run the PQ matcher of every batch inside its own step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--port', action='store_true',
                    help='--impl reference: time the C oracle port instead of baseline/_ref')
    ap.add_argument('--reference-max-steps', type=int, default=6,
                    help='--impl reference: upper bound of timed steps (seconds of CPU work each)')
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.config])
    if args.frames and args.impl == 'ours':
        WORKLOAD['B'] = args.frames
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

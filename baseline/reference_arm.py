# -*- coding: utf-8 -*-
"""The UNMODIFIED reference (TUI-NICR/nicr-multitask-scene-analysis v0.3.0), run on the host cores.

`baseline/_ref/` holds the reference exactly as `pip install --no-deps --target baseline/_ref
/root/reference` leaves it (git-ignored, shipped to the GPU box by gpurun; installed by
`__graft_entry__.build()` in the authoring container).  Its two absent third-party dependencies
(`torchmetrics`, `nicr_scene_analysis_datasets`: state bookkeeping / import hooks only, no
arithmetic of this path) are the stand-ins in `oracle/ref_stubs/`.  Nothing of the reference is
modified or re-implemented here: the step below calls its public API and stock code path,

    PanopticPostprocessing.postprocess(...)                  model/postprocessing/panoptic.py:77-316
    PanopticQuality.update(pred, target)   (its own pool)    metric/pq.py:264-303
    MeanIntersectionOverUnion.update(pred // L, sem_target)  metric/miou.py:44-56

the way `PanopticTaskHelper.validation_step` does (task_helper/panoptic.py:104-126).

Used by `bench.py` (`--impl reference`, `cpu_baseline`) and by the `-m gpu` test that compares
the CUDA path with the live reference on the GPU box.  Test / measurement infrastructure only:
the product package never imports it.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, 'baseline', '_ref')
SOURCE = '/root/reference/src'          # authoring container only
STUBS = os.path.join(ROOT, 'oracle', 'ref_stubs')


def reference_path():
    """Directory to import `nicr_mt_scene_analysis` from, or None."""
    if os.path.isdir(os.path.join(INSTALLED, 'nicr_mt_scene_analysis')):
        return INSTALLED
    if os.path.isdir(os.path.join(SOURCE, 'nicr_mt_scene_analysis')):
        return SOURCE
    return None


def available() -> bool:
    return reference_path() is not None


def load():
    """Import the reference (appended to sys.path: nothing of the environment is shadowed by the
    stubs).  Returns its entry points."""
    path = reference_path()
    if path is None:
        raise RuntimeError('the reference is neither installed under baseline/_ref nor present at '
                           '/root/reference (run `python __graft_entry__.py build` where it is)')
    for p in (STUBS, path):
        if p not in sys.path:
            sys.path.append(p)
    from nicr_mt_scene_analysis.metric import MeanIntersectionOverUnion, PanopticQuality
    from nicr_mt_scene_analysis.metric.pq import compare_and_accumulate
    from nicr_mt_scene_analysis.model.postprocessing import get_postprocessing_class
    return dict(get_postprocessing_class=get_postprocessing_class,
                PanopticQuality=PanopticQuality, MeanIntersectionOverUnion=MeanIntersectionOverUnion,
                compare_and_accumulate=compare_and_accumulate, path=path)


def build_postprocessing(ref, is_thing, has_orientation, top_k, **instance_kwargs):
    get = ref['get_postprocessing_class']
    return get('panoptic', semantic_postprocessing=get('semantic')(),
               instance_postprocessing=get('instance', top_k_instances=top_k, **instance_kwargs)(),
               semantic_classes_is_thing=is_thing,
               semantic_class_has_orientation=has_orientation)()


class ReferenceStep:
    """One pass of the hot path over a bounded sample of a workload, on the host cores."""

    def __init__(self, workload, sample_frames, threads=None, pq_workers=None, seed=1):
        import torch
        from nicr_mt_scene_analysis_b200 import testing
        w = workload
        self.w = w
        self.n = int(sample_frames)
        self.cores = int(threads or len(os.sched_getaffinity(0)))
        torch.set_num_threads(self.cores)       # torchrun exports OMP_NUM_THREADS=1
        ref = load()
        self.ref = ref
        C = w['C']
        self.L, self.offset = 1 << 16, 256 ** 3
        is_thing = testing.default_is_thing(C)
        has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
        self.post = build_postprocessing(ref, is_thing, has_ori, w['top_k'])
        # the pool of PanopticQuality (pq.py:213-218, 'spawn'): one intra-op thread per worker,
        # otherwise workers x threads oversubscribe the cores (1.7 instead
        # of 64 frames/s); the parent keeps all cores for the post-processing
        saved = os.environ.get('OMP_NUM_THREADS')
        os.environ['OMP_NUM_THREADS'] = '1'
        try:
            self.pq = ref['PanopticQuality'](
                num_categories=C + 1, ignored_label=0, max_instances_per_category=self.L,
                offset=self.offset, is_thing=[False] + list(is_thing),
                num_workers=pq_workers or min(self.cores, 32))
        finally:
            if saved is None:
                os.environ.pop('OMP_NUM_THREADS', None)
            else:
                os.environ['OMP_NUM_THREADS'] = saved
        self.miou = ref['MeanIntersectionOverUnion'](n_classes=C + 1, ignore_first_class=True)
        data = testing.make_batch(self.n, C, w['H'], w['W'], w['K'], seed=seed,
                                  with_orientation=w['ori'], quantize=None)
        self.data = data
        self.batch = testing.make_batch_dict(self.n, w['H'], w['W'])
        r = self._post()
        self.tgt_pan, self.tgt_sem = testing.make_eval_targets(
            r['panoptic_segmentation_deeplab'], self.L)

    def _post(self):
        d = self.data
        inst = (d['heat'], d['offset']) + ((d['orientation'],) if self.w['ori'] else ())
        # the reference works in place on some inputs (instance.py:86-88, panoptic.py:107)
        raw = ((d['logits'].clone(), tuple(t.clone() for t in inst)), (None, None))
        return self.post.postprocess(raw, self.batch, is_training=False)

    def step(self):
        r = self._post()
        pan = r['panoptic_segmentation_deeplab'].cpu()
        self.pq.update(pan, self.tgt_pan)
        self.miou.update(preds=(pan // self.L).cpu(), target=self.tgt_sem.cpu())
        return r

    def compute(self):
        out = self.pq.compute(suffix='_deeplab')
        out['semantic_deeplab_miou'] = self.miou.compute()
        return out

    def close(self):
        pq, self.pq = self.pq, None
        if pq is not None:
            try:
                pq.workers.terminate()
                pq.workers.join()
            except Exception:
                pass


def time_reference(workload, sample_frames, steps, warmup, threads=None):
    """-> dict(value frames/s, s_per_step, cores, frames, path).  `warmup` untimed steps first
    (the first one also absorbs the start-up of the PQ worker processes)."""
    rs = ReferenceStep(workload, sample_frames, threads=threads)
    try:
        for _ in range(max(warmup, 1)):
            rs.step()
        rs.pq.reset()
        rs.miou.reset()
        t0 = time.perf_counter()
        for _ in range(steps):
            rs.step()
        res = rs.compute()
        dt = time.perf_counter() - t0
        return dict(value=rs.n * steps / dt, s_per_step=dt / steps, cores=rs.cores,
                    frames=rs.n * steps, path=rs.ref['path'],
                    all_pq=float(res['all_deeplab_pq']), miou=float(res['semantic_deeplab_miou']))
    finally:
        rs.close()

# -*- coding: utf-8 -*-
"""Host-buffer front end: decoder outputs in (pinned) host memory in, panoptic results in
host memory out -- the shape of the reference's own path, whose panoptic outputs are CPU
tensors (model/postprocessing/panoptic.py:143-152) and whose metrics run on the CPU
(task_helper/panoptic.py:56-69).

The batch is cut into chunks of a few frames; host->device copies of chunk i+1 run on a
copy stream while chunk i is post-processed (and evaluated) on the compute stream and its
dense results are copied back, so the PCIe transfers -- the end-to-end bound of this path --
overlap the kernels.  `run()` only enqueues work and `finish()` waits for that batch alone, so
a caller can enqueue batch k+1 before building the python structures of batch k (with a second
set of `out` buffers) and keep the host->device link busy across batches.
"""
from typing import Dict, Optional

import torch

from ._results import InstanceTables
from .metric.fused import PanopticEvaluation
from .model.postprocessing.panoptic import PanopticPostprocessing


class PanopticHostPipeline:
    N_SLOTS = 2     # staging slots of `chunk_frames` frames each, used round robin across calls

    def __init__(self, postprocessing: PanopticPostprocessing,
                 evaluation: Optional[PanopticEvaluation] = None, chunk_frames: int = 8,
                 device=None):
        self.post = postprocessing
        self.evaluation = evaluation
        self.chunk = int(chunk_frames)
        self.device = torch.device(device) if device is not None else \
            torch.device('cuda', torch.cuda.current_device())
        self._copy_stream = torch.cuda.Stream(self.device)
        self._compute_stream = torch.cuda.Stream(self.device)
        self._slots = None
        self._chunks_issued = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staging(self, like: Dict[str, torch.Tensor]):
        key = tuple((k, tuple(v.shape[1:]), v.dtype) for k, v in like.items())
        if self._slots is None or self._slots[0] != key:
            slots = [{k: torch.empty((self.chunk,) + tuple(v.shape[1:]), dtype=v.dtype,
                                     device=self.device) for k, v in like.items()}
                     for _ in range(self.N_SLOTS)]
            self._slots = (key, slots, [None] * self.N_SLOTS)
            self._chunks_issued = 0
        return self._slots[1], self._slots[2]

    @staticmethod
    def pinned_like(shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    def run(self, inputs: Dict[str, torch.Tensor], batch, targets: Optional[Dict] = None,
            out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, object]:
        """inputs: host tensors 'logits' (B,C,H,W), 'heat' (B,1,H,W), 'offset' (B,2,H,W),
        optional 'orientation' (B,2,H,W); targets (optional, host): 'panoptic' (B,H,W) int64,
        'semantic' (B,H,W) uint8 -> fed to `evaluation`.  Returns host tensors
        'panoptic_segmentation_deeplab' (int64), '..._instance_idx' (uint8) and the per-frame
        python structures.  Pass pinned tensors (and `out`) to get asynchronous copies."""
        B = inputs['logits'].shape[0]
        H, W = inputs['logits'].shape[-2:]
        if B == 0:
            raise ValueError('PanopticHostPipeline.run: empty batch')
        staged = dict(inputs)
        if targets is not None:
            staged['_tgt_pan'] = targets['panoptic']
            staged['_tgt_sem'] = targets['semantic']
        slots, slot_free = self._staging(staged)
        if out is None:
            out = {'panoptic_segmentation_deeplab': self.pinned_like((B, H, W), torch.int64),
                   'panoptic_segmentation_deeplab_instance_idx':
                       self.pinned_like((B, H, W), torch.uint8)}
        pan_h = out['panoptic_segmentation_deeplab']
        inst_h = out['panoptic_segmentation_deeplab_instance_idx']
        tables = []
        h2d = d2h = 0
        # the kernels are ordered after the caller's stream (metric states may have been reset
        # there) and the caller's stream after this batch; the copy stream only moves host
        # buffers into staging slots guarded by their own events, so it is free to run ahead
        cur = torch.cuda.current_stream(self.device)
        self._compute_stream.wait_stream(cur)
        for i, lo in enumerate(range(0, B, self.chunk)):
            hi = min(lo + self.chunk, B)
            n = hi - lo
            s = self._chunks_issued % self.N_SLOTS
            self._chunks_issued += 1
            slot = slots[s]
            with torch.cuda.stream(self._copy_stream):
                if slot_free[s] is not None:
                    self._copy_stream.wait_event(slot_free[s])     # previous user is done
                for k, v in staged.items():
                    slot[k][:n].copy_(v[lo:hi], non_blocking=True)
                    h2d += v[lo:hi].numel() * v.element_size()
                ready = torch.cuda.Event()
                ready.record(self._copy_stream)
            with torch.cuda.stream(self._compute_stream):
                self._compute_stream.wait_event(ready)
                ori = slot['orientation'][:n] if 'orientation' in slot else None
                sem, inst, pan, pan_sem, tab = self.post._forward_kernels(
                    slot['logits'][:n], slot['heat'][:n], slot['offset'][:n], ori)
                if targets is not None and self.evaluation is not None:
                    # private copies (28 MB per 8 frames, nothing next to the PCIe time): the
                    # metric may re-read the targets of a frame a few updates later (large-frame
                    # path), when the staging slot already holds another chunk
                    self.evaluation.update(pan, slot['_tgt_pan'][:n].clone(),
                                           slot['_tgt_sem'][:n].clone())
                pan_h[lo:hi].copy_(pan, non_blocking=True)
                inst_h[lo:hi].copy_(inst, non_blocking=True)
                tab.prefetch()                  # per-instance tables -> pinned host memory
                d2h += pan.numel() * 8 + inst.numel() + tab.nbytes
                done = torch.cuda.Event()
                done.record(self._compute_stream)
                slot_free[s] = done
                # keep the chunk's device tensors alive until the stream has consumed them
                for t in (sem, inst, pan, pan_sem):
                    t.record_stream(self._compute_stream)
            tables.append(tab)
        cur.wait_event(done)
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        out['_tables'] = tables
        out['_done'] = done                     # everything of this batch has landed after it
        return out

    @staticmethod
    def finish(out: Dict[str, object], with_orientation: bool = True) -> Dict[str, object]:
        """Block until the copies of THIS batch have landed and build the python structures."""
        ids, meta, orientations = [], [], []
        for tab in out.pop('_tables'):
            tab: InstanceTables
            ids += tab.panoptic_ids()
            meta += tab.meta()
            if with_orientation:
                orientations += tab.orientations()
        out.pop('_done').synchronize()
        out['panoptic_segmentation_deeplab_ids'] = ids
        out['panoptic_segmentation_deeplab_instance_meta'] = meta
        if with_orientation:
            out['orientations_panoptic_segmentation_deeplab_instance'] = orientations
        return out

# -*- coding: utf-8 -*-
"""Panoptic post-processing on the GPU (API of model/postprocessing/panoptic.py:23-316).

`_postprocess_inference` issues ONE C-ABI call, `npb_panoptic_forward`, which enqueues
    centre NMS/top-k  ->  fused arg-max + offset grouping + votes + orientation sums
    ->  per-frame instance table  ->  panoptic id map
on the current CUDA stream; the packed per-instance tables reach the host in one copy when the
python dicts of the reference API are built (immediately, or on first access with
`async_results=True`).

Differences to the reference that a caller can observe (all documented in DESIGN.md):
  * dense outputs stay on the CUDA device (the reference moves the panoptic outputs to the
    CPU, panoptic.py:143-152); `.cpu()` in the callers keeps working;
  * a few bulky, rarely read entries are deferred (see _results.ResultDict);
  * more than 255 centres in a frame raise instead of silently wrapping uint8 ids, unless the
    instance post-processing was built with `on_overflow='wrap'` (the reference's result).
Extra keyword arguments (accepted through **kwargs like the reference's):
  * `async_results=True`: do not block on the table download; the dict / list entries are
    built on first access.
  * `stuff_area=n` (default 0 = off = the reference): Panoptic-DeepLab's filter, stuff segments
    of fewer than n pixels become void.
"""
from ctypes import c_float, c_int, c_int64
from typing import Tuple

import numpy as np
import torch

from ... import _lib
from ..._results import InstanceTables, ResultDict
from ...utils.fullres import fullres_key, valid_region_and_fullres_shape
from ._base import DensePostprocessingBase
from .instance import InstancePostprocessing
from .semantic import SemanticPostprocessing, widen_u8


class PanopticPostprocessing(DensePostprocessingBase):
    def __init__(
        self,
        semantic_postprocessing: SemanticPostprocessing,
        instance_postprocessing: InstancePostprocessing,
        semantic_classes_is_thing: Tuple[bool],
        semantic_class_has_orientation: Tuple[bool],
        normalized_offset: bool = True,
        compute_scores: bool = False,
        **kwargs
    ) -> None:
        super().__init__()
        self._semantic_postprocessing = semantic_postprocessing
        self._instance_postprocessing = instance_postprocessing
        # both tuples are WITHOUT void (network classes); panoptic labels are class + 1
        self._is_thing = tuple(bool(t) for t in semantic_classes_is_thing)
        self._has_orientation = tuple(bool(t) for t in semantic_class_has_orientation)
        self._thing_class_ids = np.where(self._is_thing)[0]
        self._thing_ids_panoptic = self._thing_class_ids + 1
        self._orientation_ids = np.where(self._has_orientation)[0] + 1
        self._normalized_offset = normalized_offset
        self._compute_scores = compute_scores
        self._max_instances_per_category = 1 << 16
        self._async_results = bool(kwargs.get('async_results', False))
        # Panoptic-DeepLab's `stuff_area`: stuff segments of fewer pixels become void.  NOT in the
        # reference (its merge keeps every stuff class present, panoptic_merge.py:213-223), so the
        # default 0 = off is the reference's result; see npb_filter_stuff_area.
        self._stuff_area = int(kwargs.get('stuff_area', 0))
        if self._stuff_area < 0:
            raise ValueError('stuff_area must be >= 0')
        self._ws = {}
        self._fused_evaluation = None       # see fuse_evaluation()

    @property
    def max_instances_per_category(self):
        return self._max_instances_per_category

    def _postprocess_training(self, data, batch):
        (s_output, i_output), (s_side_outputs, i_side_outputs) = data
        r_sem = self._semantic_postprocessing._postprocess_training((s_output, s_side_outputs), batch)
        r_ins = self._instance_postprocessing._postprocess_training((i_output, i_side_outputs), batch)
        return {**r_sem, **r_ins}

    # ------------------------------------------------------------------ fused kernel chain
    def _forward_plan(self, dev, B, C, H, W):
        """Everything of a `_forward_kernels` call that only depends on (device, stream, shape,
        configuration): scratch workspace, layout of the output allocation, the constant part of
        the C argument list.  Built once per shape -- the per-call host work is one allocation,
        a dozen integer additions and the call."""
        post = self._instance_postprocessing
        ks = post._heatmap_nms_kernel_size
        use_thr = post._offset_distance_threshold is not None
        key = (dev, torch.cuda.current_stream(dev).cuda_stream, B, C, H, W, ks,
               post._heatmap_threshold, post._top_k_instances, post._heatmap_apply_foreground_mask,
               self._normalized_offset, post._offset_distance_threshold)
        plan = self._ws.get(key)
        if plan is not None:
            return plan
        if len(self._is_thing) != C:
            raise ValueError(f'semantic_classes_is_thing has {len(self._is_thing)} entries, the '
                             f'logits have {C} classes')
        L = _lib.lib()
        # scratch (candidate lists, vote histograms, orientation sums) is reused across calls:
        # the calls are ordered on the stream, nothing in it outlives a call
        # (one scratch buffer per stream: calls on different streams may run concurrently)
        # (npb_panoptic_forward_workspace_init once per buffer and shape: the chain itself contains
        # no memset and leaves the workspace ready for the next call)
        ws_bytes = L.npb_panoptic_forward_workspace_bytes(B, C, H, W, ks)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.npb_panoptic_forward_workspace_init(
            _lib.ptr(ws), c_int(B), c_int(C), c_int(H), c_int(W), c_int(ks),
            _lib.stream_ptr(dev)), 'npb_panoptic_forward_workspace_init')
        # all outputs of a call live in ONE fresh allocation: [pan i64 | tables | sem | inst | pan_sem]
        P = H * W
        offsets, tab_bytes = InstanceTables.layout(B)
        pan_bytes = B * P * 8
        u8_bytes = (B * P + 15) // 16 * 16
        o = pan_bytes + tab_bytes
        plan = dict(
            ws=ws, P=P, pan_bytes=pan_bytes, tab_bytes=tab_bytes, u8_bytes=u8_bytes,
            total=pan_bytes + tab_bytes + 3 * u8_bytes, sem_off=o, inst_off=o + u8_bytes,
            pan_sem_off=o + 2 * u8_bytes,
            table_offs=tuple(pan_bytes + offsets[name][0] for name in (
                'centers_yx', 'n_centers', 'center_score', 'inst_class', 'inst_pan_id', 'inst_area',
                'inst_angle', 'status')),
            shape_args=(c_int(B), c_int(C), c_int(H), c_int(W),
                        _lib.host_lut(self._is_thing, C), _lib.host_lut(self._has_orientation, C),
                        c_float(post._heatmap_threshold), c_int(ks), c_int(post._top_k_instances),
                        c_int(int(post._heatmap_apply_foreground_mask)),
                        c_int(int(self._normalized_offset)), c_int(int(use_thr)),
                        c_float(float(post._offset_distance_threshold) if use_thr else 0.0),
                        c_int64(self._max_instances_per_category), ws.data_ptr()))
        if len(self._ws) >= 8:      # shapes rarely change; do not hoard scratch if they do
            self._ws.clear()
        self._ws[key] = plan
        return plan

    def _forward_kernels(self, logits, heat, offset, orientation, eval_args=None, pipeline=None):
        logits = _lib.require_cuda(logits, 'semantic logits', torch.float32, 4)
        dev = logits.device
        heat = _lib.require_cuda(heat, 'center_heatmap', torch.float32, 4)
        offset = _lib.require_cuda(offset, 'center_offset', torch.float32, 4)
        if orientation is not None:
            orientation = _lib.require_cuda(orientation, 'orientation', torch.float32, 4)
        B, C, H, W = logits.shape
        if heat.shape != (B, 1, H, W) or offset.shape != (B, 2, H, W):
            raise ValueError('instance outputs do not match the semantic logits in shape')
        L = _lib.lib()
        plan = self._forward_plan(dev, B, C, H, W)
        P, pan_bytes, tab_bytes = plan['P'], plan['pan_bytes'], plan['tab_bytes']
        buf = torch.empty(plan['total'], dtype=torch.uint8, device=dev)
        base = buf.data_ptr()
        pan = buf[:pan_bytes].view(torch.int64).view(B, H, W)
        tables = InstanceTables(B, dev, storage=buf[pan_bytes:pan_bytes + tab_bytes])
        o = plan['sem_off']
        sem = buf[o:o + B * P].view(B, H, W)
        o = plan['inst_off']
        inst = buf[o:o + B * P].view(B, H, W)
        o = plan['pan_sem_off']
        pan_sem = buf[o:o + B * P].view(B, H, W)
        args = (logits.data_ptr(), heat.data_ptr(), offset.data_ptr(),
                None if orientation is None else orientation.data_ptr()) + plan['shape_args'] + \
            (base + plan['sem_off'], base + plan['inst_off'], base, base + plan['pan_sem_off']) + \
            tuple(base + t for t in plan['table_offs'])
        stream = torch.cuda.current_stream(dev).cuda_stream
        with _lib.on_device(dev):
            if eval_args is None:
                rc = L.npb_panoptic_forward(*args, stream)
                where = 'npb_panoptic_forward'
            elif pipeline is None:       # the last stage also evaluates the ids it writes
                import ctypes
                rc = L.npb_panoptic_forward_eval(*args, ctypes.byref(eval_args), stream)
                where = 'npb_panoptic_forward_eval'
            else:
                # ... and leaves its matcher to the next call, which runs it next to its own centre
                # detection + grouping (`pipeline` = (pending npb_eval_args or None, its batch size))
                import ctypes
                pending, pending_B = pipeline
                rc = L.npb_panoptic_forward_eval_pipelined(
                    *args, ctypes.byref(eval_args),
                    ctypes.byref(pending) if pending is not None else None, c_int(pending_B), stream)
                where = 'npb_panoptic_forward_eval_pipelined'
        _lib.raise_for_code(rc, where)
        post = self._instance_postprocessing
        for b in post._overflow_frames(tables):         # on_overflow='wrap' only (one sync)
            self._redo_wrapped_frame(plan, tables, b, heat, offset, orientation, sem, inst, pan,
                                     pan_sem, C, H, W)
        if self._stuff_area > 0:
            counts = torch.empty((B, C + 1), dtype=torch.int32, device=dev)
            _lib.check(L.npb_filter_stuff_area(
                _lib.ptr(pan), _lib.ptr(pan_sem), c_int(B), c_int64(P), c_int(C + 1),
                c_int64(self._max_instances_per_category), c_int64(self._stuff_area), c_int64(0),
                _lib.ptr(counts), _lib.stream_ptr(dev)), 'npb_filter_stuff_area')
        return sem, inst, pan, pan_sem, tables

    def _redo_wrapped_frame(self, plan, tables, b, heat, offset, orientation, sem, inst, pan,
                            pan_sem, C, H, W) -> None:
        """A frame with more than 255 centres, redone like the reference computes it
        (instance.py:231-236 wraps the uint8 ids): complete centre list from the workspace of
        the call, grouping with all centres (ids mod 256), then the ordinary instance table and
        id writer on the frame's rows."""
        post = self._instance_postprocessing
        dev = heat.device
        L = _lib.lib()
        centers = post._wide_centers(tables, plan['ws'], heat, b)
        hist = torch.empty((_lib.MAX_INST, C), dtype=torch.int32, device=dev)
        ori_sum = None if orientation is None else \
            torch.empty((_lib.MAX_INST, 2), dtype=torch.float64, device=dev)
        thing_lut = _lib.host_lut(self._is_thing, C)
        post._regroup_wide(tables, b, centers, sem[b], None, offset[b],
                           None if orientation is None else orientation[b], C, H, W, thing_lut,
                           self._normalized_offset, inst[b], hist, ori_sum)
        stream = _lib.stream_ptr(dev)
        _lib.check(L.npb_finalize_instances(
            _lib.ptr(hist), _lib.ptr(ori_sum), tables.dptr_row('n_centers', b), c_int(1), c_int(C),
            c_int(1), c_int64(self._max_instances_per_category), c_int64(0),
            _lib.host_lut(self._has_orientation, C), tables.dptr_row('inst_class', b),
            tables.dptr_row('inst_pan_id', b), tables.dptr_row('inst_area', b),
            tables.dptr_row('inst_angle', b), stream), 'npb_finalize_instances')
        _lib.check(L.npb_write_panoptic(
            _lib.ptr(sem[b]), _lib.ptr(inst[b]), tables.dptr_row('inst_pan_id', b),
            tables.dptr_row('inst_class', b), c_int(1), c_int(C), c_int(H), c_int(W), thing_lut,
            c_int64(self._max_instances_per_category), _lib.ptr(pan[b]), _lib.ptr(pan_sem[b]),
            _lib.stream_ptr(dev)), 'npb_write_panoptic')

    # ------------------------------------------------------------------ fused evaluation
    PIPELINE_MAX_FRAMES = 16    # largest batch whose matcher is pipelined over consecutive calls

    def fuse_evaluation(self, evaluation, pipeline_matching: bool = False) -> None:
        """Attach a `metric.PanopticEvaluation` (or None to detach).  When the batch handed to
        `postprocess` carries the ground truth (`panoptic_fullres`, `semantic_fullres`) at the
        resolution of the network outputs, the kernel that writes the panoptic ids also feeds
        PQ and mIoU (task_helper/panoptic.py:104-126) -- the ids are not read back.  The result
        dict then carries `_panoptic_evaluation_fused` (and `_panoptic_matches` when the batch
        has orientations) so that the task helper skips its own update.

        `pipeline_matching`: a validation LOOP.  The PQ matcher of a batch (one CTA per frame of
        pure latency) is left pending and runs next to the centre detection + grouping of the
        NEXT call, or when anything reads / resets the PQ states (`compute`, `check_status`,
        `reset`, a stand-alone `update`).  States and results are identical; only batches that
        need their matches right away (orientation MAAE) are matched in their own call."""
        if self._fused_evaluation is not None and evaluation is not self._fused_evaluation:
            self._fused_evaluation.pq._flush_deferred()
        if evaluation is not None and self._instance_postprocessing._on_overflow == 'wrap':
            raise ValueError("fuse_evaluation: on_overflow='wrap' redoes frames after the call, "
                             'when the fused evaluation has already counted them; update the '
                             'metrics from the result dict instead')
        if evaluation is not None and self._stuff_area > 0:
            raise ValueError('fuse_evaluation: the stuff-area filter rewrites the ids after the '
                             'kernel that writes (and would evaluate) them')
        if evaluation is not None:
            # the fused kernels evaluate with THIS object's id geometry and class count
            if evaluation.pq.max_instances_per_category != self._max_instances_per_category:
                raise ValueError('fuse_evaluation: the PanopticQuality object uses '
                                 f'{evaluation.pq.max_instances_per_category} instances per '
                                 f'category, the post-processing {self._max_instances_per_category}')
            if evaluation.pq.num_categories != len(self._is_thing) + 1 or \
                    evaluation.miou._n_classes != len(self._is_thing) + 1:
                raise ValueError('fuse_evaluation: the metrics must count the network classes + void')
        self._fused_evaluation = evaluation
        self._pipeline_matching = bool(pipeline_matching) and evaluation is not None

    def _fused_eval_args(self, batch, logits):
        ev = getattr(self, '_fused_evaluation', None)
        if ev is None:
            return None
        pan_t, sem_t = batch.get(fullres_key('panoptic')), batch.get(fullres_key('semantic'))
        if not isinstance(pan_t, torch.Tensor) or not isinstance(sem_t, torch.Tensor):
            return None
        shape = (logits.shape[0],) + tuple(logits.shape[-2:])
        if tuple(pan_t.shape) != shape or tuple(sem_t.shape) != shape:
            return None         # evaluation happens at dataset resolution: not the same maps
        want_matches = 'orientations_present' in batch and hasattr(ev.pq, 'update_mae')
        # Every matcher CTA owns a whole SM while it works through its latency-bound phases:
        # next to the (HBM-saturating) grouping kernel of a large batch that costs more than the
        # matcher's place at the end of the step (measured: 64 frames 530x730, 800 -> 850 us),
        # for a few frames it removes it from the step (8 frames 480x640, 108 -> 99 us)
        pipelined = getattr(self, '_pipeline_matching', False) and not want_matches and \
            shape[0] <= self.PIPELINE_MAX_FRAMES
        if not pipelined:
            ev.pq._flush_deferred()         # the call below clears the hand-over workspace
        return ev.eval_args(pan_t, sem_t, want_matches=want_matches) + (ev, pipelined)

    def _thing_mask(self, sem_u8: torch.Tensor) -> torch.Tensor:
        """panoptic.py:123-127 `isin(semantic idx, thing ids)` -> bool (B,H,W)."""
        out = torch.empty(sem_u8.shape, dtype=torch.uint8, device=sem_u8.device)
        C = len(self._is_thing)
        _lib.check(_lib.lib().npb_thing_mask(
            _lib.ptr(sem_u8), c_int64(sem_u8.numel()), c_int(C), _lib.host_lut(self._is_thing, C),
            _lib.ptr(out), _lib.stream_ptr(sem_u8.device)), 'npb_thing_mask')
        return out.view(torch.bool)

    def _postprocess_inference(self, data, batch):
        (s_output, i_output), (s_side_outputs, i_side_outputs) = data
        with_orientation = (3 == len(i_output))
        center_heatmap, center_offset = i_output[0], i_output[1]
        orientation = i_output[2] if with_orientation else None

        fused = self._fused_eval_args(batch, s_output)
        pipeline = None
        if fused and fused[3]:
            pq = fused[2].pq
            capturing = torch.cuda.is_current_stream_capturing()
            B = s_output.shape[0]
            if capturing and pq._deferred is not None and pq._deferred['B'] != B:
                raise RuntimeError('a pipelined update of another batch size is pending: call '
                                   'check_status() / compute() before capturing this step')
            if capturing and pq._deferred is not None:
                # every replay runs the matcher of the replay before it: same buffers, same
                # (shared) status words as this call -- the descriptor of this call itself
                pipeline = (fused[0], B)
            elif not capturing:
                d = pq._deferred
                pipeline = (d['args'], d['B']) if d is not None else (None, 0)
            # (a capture that finds nothing pending cannot start a pipeline: it would replay a
            # pixel pass whose matcher never runs -- that call is matched in place)
        sem, inst, pan, pan_sem, tables = self._forward_kernels(
            s_output, center_heatmap, center_offset, orientation,
            eval_args=fused[0] if fused else None, pipeline=pipeline)

        # semantic + instance entries (panoptic.py:86-94); the class map is shared
        r = ResultDict(semantic_output=s_output, semantic_side_outputs=s_side_outputs)
        self._semantic_postprocessing._fill_inference_entries(r, s_output, batch, sem_u8=sem)
        r.update(self._instance_postprocessing._postprocess_inference((i_output, i_side_outputs),
                                                                     batch))

        # panoptic entries (panoptic.py:128-167)
        r.defer('panoptic_foreground_mask', lambda: self._thing_mask(sem))
        r['panoptic_segmentation_deeplab'] = pan
        r['_panoptic_segmentation_deeplab_semantic_idx_u8'] = pan_sem
        r.defer('panoptic_segmentation_deeplab_semantic_idx', lambda: widen_u8(pan_sem))
        r['panoptic_segmentation_deeplab_instance_idx'] = inst
        r['_panoptic_instance_tables'] = tables
        if fused and pipeline is not None:
            pq = fused[2].pq
            captured = torch.cuda.is_current_stream_capturing()
            pq._deferred_matcher_issued(captured=captured)     # the pending one has been enqueued
            pq._set_deferred(fused[0], fused[1], pan, s_output.shape[0])
            # a CUDA-graph replay leaves the same update pending again (graph.CapturedStep)
            r['_panoptic_evaluation_pipelined'] = \
                lambda pq=pq, a=fused[0], k=fused[1], B=s_output.shape[0]: pq._set_deferred(a, k, None, B)
        elif fused:
            fused[2].pq._fused_issued(fused[1], pan)
        if fused:
            # PQ / mIoU states already hold this batch (or will, before anybody can read them);
            # a task helper must not add it again
            r['_panoptic_evaluation_fused'] = True
            if fused[1].get('matches') is not None:
                r['_panoptic_matches'] = (fused[1]['matches'], fused[1]['n_matches'])
        if self._async_results:
            r.defer('panoptic_segmentation_deeplab_ids', tables.panoptic_ids)
            r.defer_with_dict('panoptic_segmentation_deeplab_instance_meta',
                              lambda r_: self._meta(r_, tables))
        else:
            r['panoptic_segmentation_deeplab_ids'] = tables.panoptic_ids()
            r['panoptic_segmentation_deeplab_instance_meta'] = tables.meta(with_orientation)

        if self._compute_scores:
            self._add_scores(r, s_output, pan, pan_sem, tables)

        # full resolution (panoptic.py:242-291)
        crop, shape = valid_region_and_fullres_shape(batch, 'instance')
        dense = ['panoptic_segmentation_deeplab', 'panoptic_segmentation_deeplab_instance_idx',
                 'panoptic_segmentation_deeplab_semantic_idx']
        if self._compute_scores:
            dense += [f'panoptic_segmentation_deeplab_{k}_score'
                      for k in ('semantic', 'instance', 'panoptic')]
        identity = self._is_identity_resize((pan.shape[-2], pan.shape[-1]), crop, shape)
        for key in dense:
            if identity:
                r.alias(fullres_key(key), key)
            else:
                r.defer_with_dict(
                    fullres_key(key),
                    lambda r_, key=key: self._crop_to_valid_region_and_resize_prediction(
                        r_[key], crop, shape, mode='nearest'))

        # orientation (panoptic.py:294-314)
        if with_orientation:
            if self._async_results:
                r.defer('orientations_panoptic_segmentation_deeplab_instance', tables.orientations)
            else:
                # (the 'orientation' fields of the meta dicts were filled from the same table)
                r['orientations_panoptic_segmentation_deeplab_instance'] = tables.orientations()
        return r

    @staticmethod
    def _meta(r: ResultDict, tables: InstanceTables):
        return tables.meta('orientations_panoptic_segmentation_deeplab_instance' in r)

    # ------------------------------------------------------------------ optional score maps
    def _add_scores(self, r, logits, pan, pan_sem, tables):
        """panoptic.py:171-239 (`compute_scores=True`): dense semantic / instance / panoptic
        score maps from `npb_panoptic_scores` (csrc/scores.cu) and the per-instance score
        fields of the meta dicts."""
        logits = _lib.require_cuda(logits, 'semantic logits', torch.float32, 4)
        B, C, H, W = logits.shape
        dev = logits.device
        inst = r['panoptic_segmentation_deeplab_instance_idx']
        maps = torch.empty((3, B, H, W), dtype=torch.float32, device=dev)
        per_inst = torch.empty((2, B, _lib.MAX_INST), dtype=torch.float32, device=dev)
        inst_sum = torch.empty((B, _lib.MAX_INST), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().npb_panoptic_scores(
            _lib.ptr(logits), _lib.ptr(pan_sem), _lib.ptr(inst), tables.dptr('inst_class'),
            tables.dptr('inst_area'), tables.dptr('center_score'), c_int(B), c_int(C), c_int(H),
            c_int(W), _lib.ptr(inst_sum), _lib.ptr(maps[0]), _lib.ptr(maps[1]), _lib.ptr(maps[2]),
            _lib.ptr(per_inst[0]), _lib.ptr(per_inst[1]), _lib.stream_ptr(dev)),
            'npb_panoptic_scores')
        r['panoptic_segmentation_deeplab_semantic_score'] = maps[0]
        r['panoptic_segmentation_deeplab_instance_score'] = maps[1]
        r['panoptic_segmentation_deeplab_panoptic_score'] = maps[2]

        def fill(meta):
            mean_sem, pan_score = (x.tolist() for x in per_inst.cpu())
            cls = tables['inst_class']
            for b, ids_b in enumerate(tables.panoptic_ids()):
                for pan_id, ins_id in ids_b.items():
                    entry = meta[b][ins_id]
                    entry['semantic_score'] = mean_sem[b][ins_id]
                    entry['semantic_idx'] = int(cls[b, ins_id])
                    entry['panoptic_score'] = pan_score[b][ins_id]
                    entry['panoptic_id'] = pan_id
            return meta

        key = 'panoptic_segmentation_deeplab_instance_meta'
        if r.is_deferred(key):
            r.defer_with_dict(key, lambda r_: fill(self._meta(r_, tables)))
        else:
            fill(r[key])

# -*- coding: utf-8 -*-
"""Instance post-processing on the GPU (API of model/postprocessing/instance.py:24-468).

Constructor arguments, the three `_get_*` methods that `PanopticPostprocessing` and the
reference's tests call directly, and the inference result keys are those of the reference;
the work is done by csrc/centers.cu (NMS + top-k), csrc/group.cu (offset grouping) and
csrc/misc.cu (orientation averaging) through the C ABI.
"""
from ctypes import c_float, c_int, c_int64
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from ... import _lib
from ..._results import InstanceTables, ResultDict
from ...utils.fullres import fullres_key, valid_region_and_fullres_shape
from ._base import DensePostprocessingBase


class InstancePostprocessing(DensePostprocessingBase):
    def __init__(
        self,
        heatmap_threshold: float = 0.1,
        heatmap_nms_kernel_size: int = 3,
        heatmap_apply_foreground_mask: bool = False,
        top_k_instances: int = 64,
        normalized_offset: bool = True,
        offset_distance_threshold: Union[None, int] = None,
        **kwargs
    ) -> None:
        super().__init__()
        assert heatmap_nms_kernel_size % 2 == 1
        assert top_k_instances <= 254
        self._heatmap_threshold = heatmap_threshold
        self._heatmap_nms_kernel_size = heatmap_nms_kernel_size
        self._heatmap_nms_padding = (heatmap_nms_kernel_size - 1) // 2
        self._heatmap_apply_foreground_mask = heatmap_apply_foreground_mask
        self._top_k_instances = top_k_instances
        self._normalized_offset = normalized_offset
        self._offset_distance_threshold = offset_distance_threshold
        self.debug = kwargs.get('debug', False)
        # More than 255 centres in a frame (k-th-value ties of a saturated heat-map).  'raise'
        # (default): NpbError(TOO_MANY_CENTERS) when the results are read.  'wrap': the frame is
        # redone the way the reference computes it -- uint8 ids that wrap silently
        # (instance.py:236: centre 256 is "no instance", centre 257 joins instance 1, the meta
        # dict keeps all centres) -- at the price of one status read-back per call.
        self._on_overflow = kwargs.get('on_overflow', 'raise')
        if self._on_overflow not in ('raise', 'wrap'):
            raise ValueError("on_overflow must be 'raise' or 'wrap'")

    # ------------------------------------------------------------------ kernels
    def _run_centers(self, heat: torch.Tensor, fg_u8: Optional[torch.Tensor]) -> InstanceTables:
        """npb_instance_centers -> tables with centres / counts / scores (on device)."""
        heat = _lib.require_cuda(heat, 'center_heatmap', torch.float32, 4)
        B, one, H, W = heat.shape
        assert one == 1
        dev = heat.device
        L = _lib.lib()
        tables = InstanceTables(B, dev)
        ws = torch.empty(L.npb_instance_centers_workspace_bytes(B, H, W, self._heatmap_nms_kernel_size),
                         dtype=torch.uint8, device=dev)
        apply_fg = bool(self._heatmap_apply_foreground_mask)
        if apply_fg and fg_u8 is None:
            raise ValueError('heatmap_apply_foreground_mask=True needs a foreground mask')
        _lib.check(L.npb_instance_centers(
            _lib.ptr(heat), c_int(B), c_int(H), c_int(W), c_float(self._heatmap_threshold),
            c_int(self._heatmap_nms_kernel_size), c_int(self._top_k_instances),
            _lib.ptr(fg_u8) if apply_fg else None, c_int(int(apply_fg)), _lib.ptr(ws),
            tables.dptr('centers_yx'), tables.dptr('n_centers'), tables.dptr('center_score'),
            tables.dptr('status'), _lib.stream_ptr(dev)), 'npb_instance_centers')
        tables._where = 'instance centres'
        tables._centers_ws = ws         # holds the complete centre list of an overflowing frame
        return tables

    # ------------------------------------------------------------------ > 255 centres
    def _overflow_frames(self, tables: InstanceTables) -> List[int]:
        """Frames of the call that reported more than 255 centres (one synchronous read of the
        status words; only with `on_overflow='wrap'`)."""
        if self._on_overflow != 'wrap':
            return []
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("on_overflow='wrap' reads the status words of every call on the "
                               'host: such a step cannot be captured into a CUDA graph')
        codes = tables.dview('status').cpu().tolist()
        return [b for b, c in enumerate(codes) if c == _lib.ERR_TOO_MANY_CENTERS]

    def _wide_centers(self, tables: InstanceTables, workspace: torch.Tensor, heat: torch.Tensor,
                      b: int):
        """npb_overflow_centers: the complete centre list of frame `b` -> device (count [1],
        centres (cap,2), scores (cap)); the host copy goes into `tables.wide[b]`."""
        B, _, H, W = heat.shape
        dev = heat.device
        cap = _lib.MAX_WIDE_CENTERS
        n_dev = torch.empty(1, dtype=torch.int32, device=dev)
        cyx = torch.empty((cap, 2), dtype=torch.int32, device=dev)
        score = torch.empty(cap, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().npb_overflow_centers(
            _lib.ptr(workspace), _lib.ptr(heat), c_int(B), c_int(H), c_int(W),
            c_int(self._heatmap_nms_kernel_size), c_int(b), _lib.ptr(n_dev), _lib.ptr(cyx),
            _lib.ptr(score), c_int(cap), _lib.stream_ptr(dev)), 'npb_overflow_centers')
        n = int(n_dev.item())
        if n < 0:
            raise _lib.NpbError(_lib.ERR_TOO_MANY_CENTERS,
                                f'more than {cap} instance centres in frame {b}')
        tables.wide[b] = (cyx[:n].cpu().numpy(), score[:n].cpu().numpy())
        return n_dev, cyx, score

    def _regroup_wide(self, tables: InstanceTables, b: int, centers, sem_b, fg_b, offset_b,
                      orientation_b, C: int, H: int, W: int, thing_lut, normalized: bool,
                      inst_b: torch.Tensor, hist: torch.Tensor, ori_sum) -> None:
        """npb_group_pixels_wide for frame `b`: ids wrapped like instance.py:236; leaves the
        frame's vote histogram / orientation sums in `hist` / `ori_sum`, puts min(n, 255) into
        the frame's `n_centers` word and its status word back to OK."""
        n_dev, cyx, _ = centers
        use_thr = self._offset_distance_threshold is not None
        dev = inst_b.device
        _lib.check(_lib.lib().npb_group_pixels_wide(
            _lib.ptr(sem_b), _lib.ptr(fg_b), _lib.ptr(offset_b), _lib.ptr(orientation_b),
            c_int(C), c_int(H), c_int(W), thing_lut, _lib.ptr(cyx), _lib.ptr(n_dev),
            c_int(int(normalized)), c_int(int(use_thr)),
            c_float(float(self._offset_distance_threshold) if use_thr else 0.0),
            _lib.ptr(inst_b), _lib.ptr(hist), _lib.ptr(ori_sum), tables.dptr_row('n_centers', b),
            tables.dptr_row('status', b), _lib.stream_ptr(dev)), 'npb_group_pixels_wide')

    @staticmethod
    def _as_fg_u8(foreground_mask: torch.Tensor, device) -> torch.Tensor:
        fg = foreground_mask.to(device)
        if fg.ndim == 4:
            fg = fg[:, 0]
        if fg.dtype == torch.bool:
            fg = fg.contiguous().view(torch.uint8)
        elif fg.dtype != torch.uint8:
            fg = (fg != 0).view(torch.uint8)
        return fg.contiguous()

    # ------------------------------------------------------------------ reference API
    def _get_instance_centers(
        self,
        center_heatmap: torch.Tensor,
        foreground_mask: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """instance.py:78-168 -> (bool (B,H,W) centre map, list of (n,2) int32 (y,x))."""
        fg = None
        if self._heatmap_apply_foreground_mask and foreground_mask is not None:
            fg = self._as_fg_u8(foreground_mask, center_heatmap.device)
        tables = self._run_centers(center_heatmap, fg)
        for b in self._overflow_frames(tables):         # on_overflow='wrap': all of them
            self._wide_centers(tables, tables._centers_ws, center_heatmap, b)
            tables.dview('status')[b] = 0
        centers = tables.centers_list()
        B, _, H, W = center_heatmap.shape
        mask = torch.zeros((B, H, W), dtype=torch.bool)
        for b, c in enumerate(centers):
            if len(c):
                mask[b, c[:, 0].long(), c[:, 1].long()] = True
        return mask.to(center_heatmap.device), centers

    def _group(self, tables: InstanceTables, center_offset: torch.Tensor, fg_u8: torch.Tensor,
               normalized: bool = False, heat: Optional[torch.Tensor] = None):
        """npb_group_pixels with an explicit foreground mask (no classes) +
        npb_finalize_instances for the areas."""
        off = _lib.require_cuda(center_offset, 'center_offset', torch.float32, 4)
        B, two, H, W = off.shape
        assert two == 2
        dev = off.device
        L = _lib.lib()
        inst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        hist = torch.empty((B, _lib.MAX_INST, 1), dtype=torch.int32, device=dev)
        use_thr = self._offset_distance_threshold is not None
        _lib.check(L.npb_group_pixels(
            None, None, _lib.ptr(fg_u8), _lib.ptr(off), None, c_int(B), c_int(1), c_int(H),
            c_int(W), None, tables.dptr('centers_yx'), tables.dptr('n_centers'),
            c_int(int(normalized)), c_int(int(use_thr)),
            c_float(float(self._offset_distance_threshold) if use_thr else 0.0), None,
            _lib.ptr(inst), _lib.ptr(hist), None, _lib.stream_ptr(dev)), 'npb_group_pixels')
        pan_dummy = torch.empty((B, _lib.MAX_INST), dtype=torch.int64, device=dev)
        _lib.check(L.npb_finalize_instances(
            _lib.ptr(hist), None, tables.dptr('n_centers'), c_int(B), c_int(1), c_int(1),
            c_int64(1 << 16), c_int64(0), None, tables.dptr('inst_class'), _lib.ptr(pan_dummy),
            tables.dptr('inst_area'), tables.dptr('inst_angle'), _lib.stream_ptr(dev)),
            'npb_finalize_instances')
        for b in self._overflow_frames(tables):         # on_overflow='wrap'
            centers = self._wide_centers(tables, tables._centers_ws, heat, b)
            self._regroup_wide(tables, b, centers, None, fg_u8[b], off[b], None, 1, H, W, None,
                               normalized, inst[b], hist[b], None)
            _lib.check(L.npb_finalize_instances(
                _lib.ptr(hist[b]), None, tables.dptr_row('n_centers', b), c_int(1), c_int(1),
                c_int(1), c_int64(1 << 16), c_int64(0), None, tables.dptr_row('inst_class', b),
                _lib.ptr(pan_dummy[b]), tables.dptr_row('inst_area', b),
                tables.dptr_row('inst_angle', b), _lib.stream_ptr(dev)), 'npb_finalize_instances')
        return inst

    def _get_instance_segmentation(
        self,
        center_heatmap: torch.Tensor,
        center_offset: torch.Tensor,
        foreground_mask: torch.Tensor
    ) -> Tuple[torch.Tensor, List[Dict[int, Dict[str, Any]]]]:
        """instance.py:170-268 -> (uint8 (B,H,W) instance ids, meta dicts).
        `center_offset` in pixels, `foreground_mask` bool (B,H,W) or (B,1,H,W)."""
        return self._segment(center_heatmap, center_offset, foreground_mask, normalized=False)

    def _segment(self, center_heatmap, center_offset, foreground_mask, normalized: bool):
        """`normalized=True`: `center_offset` is still divided by (h, w); the kernel applies
        the de-normalising multiply of instance.py:361-367 itself (same single f32 mul)."""
        dev = center_heatmap.device
        fg = self._as_fg_u8(foreground_mask, dev)
        tables = self._run_centers(center_heatmap, fg)
        heat = _lib.require_cuda(center_heatmap, 'center_heatmap', torch.float32, 4)
        inst = self._group(tables, center_offset, fg, normalized, heat)
        return inst, tables.meta()

    def _get_instance_orientation(
        self,
        orientation: torch.Tensor,
        instance_segmentation: torch.Tensor,
        foreground_mask: Optional[torch.Tensor]
    ) -> List[Dict[int, float]]:
        """instance.py:270-319 -> per frame {instance id: mean angle in rad}."""
        ori = _lib.require_cuda(orientation, 'orientation', torch.float32, 4)
        dev = ori.device
        seg = instance_segmentation.to(dev)
        if seg.ndim == 4:
            seg = seg[:, 0]
        if seg.dtype not in (torch.uint8, torch.int16, torch.int32, torch.int64):
            seg = seg.to(torch.int64)
        seg = seg.contiguous()
        B = ori.shape[0]
        P = ori.shape[-2] * ori.shape[-1]
        mask = None if foreground_mask is None else self._as_fg_u8(foreground_mask, dev)
        max_id = max(int(seg.max()) if seg.numel() else 0, 1)
        count = torch.empty((B, max_id + 1), dtype=torch.int32, device=dev)
        angle = torch.empty((B, max_id + 1), dtype=torch.float32, device=dev)
        sums = torch.empty((B, max_id + 1, 2), dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().npb_instance_orientation(
            _lib.ptr(ori), _lib.ptr(seg), c_int(_lib.dtype_code(seg)), _lib.ptr(mask), c_int(B),
            c_int64(P), c_int(max_id), _lib.ptr(count), _lib.ptr(angle), _lib.ptr(sums),
            _lib.ptr(status), _lib.stream_ptr(dev)), 'npb_instance_orientation')
        # python lists (float32 values as python floats): tensor elements read one by one
        # cost microseconds each
        count_h, angle_h = count.cpu().tolist(), angle.cpu().tolist()
        _lib.raise_for_status(status.cpu().tolist(), 'instance orientation')
        return [{i: angle_b[i] for i in range(1, max_id + 1) if count_b[i] > 0}
                for count_b, angle_b in zip(count_h, angle_h)]

    # ------------------------------------------------------------------ postprocess
    def _postprocess_training(self, data, batch):
        output, side_outputs = data
        return {'instance_output': output, 'instance_side_outputs': side_outputs}

    def _postprocess_inference(self, data, batch):
        output, side_outputs = data
        with_orientation = (3 == len(output))
        center_heatmap, center_offset = output[0], output[1]
        orientation = output[2] if with_orientation else None

        r = ResultDict(instance_output=output, instance_side_outputs=side_outputs,
                       instance_centers=center_heatmap, instance_offsets=center_offset)
        if with_orientation:
            r['instance_orientation'] = orientation

        def segment(foreground_mask, key):
            seg, meta = self._segment(center_heatmap, center_offset, foreground_mask,
                                      normalized=self._normalized_offset)
            r[key] = seg
            crop, shape = valid_region_and_fullres_shape(batch, 'instance')
            r[fullres_key(key)] = self._crop_to_valid_region_and_resize_prediction(
                seg, crop, shape, mode='nearest')
            return meta

        # i-1: ground-truth foreground (dataset evaluation), instance.py:371-397
        if 'instance_foreground' in batch:
            r['instance_segmentation_gt_meta'] = segment(batch['instance_foreground'],
                                                         'instance_segmentation_gt_foreground')
        # i-2: everything is foreground (debugging), instance.py:399-420
        if self.debug:
            segment(torch.ones_like(center_heatmap, dtype=torch.bool),
                    'instance_segmentation_all_foreground')
        if not with_orientation:
            return r

        # o-1 .. o-4, instance.py:428-466
        if all(k in batch for k in ('instance', 'orientation_foreground')):
            r['orientations_gt_instance_gt_orientation_foreground'] = \
                self._get_instance_orientation(orientation, batch['instance'],
                                               batch['orientation_foreground'])
        if all(k in batch for k in ('instance_foreground', 'orientation_foreground')):
            r['orientations_instance_segmentation_gt_orientation_foreground'] = \
                self._get_instance_orientation(orientation,
                                               r['instance_segmentation_gt_foreground'],
                                               batch['orientation_foreground'])
        if self.debug:
            r['orientations_gt_instance'] = self._get_instance_orientation(
                orientation, batch['instance'], None)
            r['orientations_instance_segmentation'] = self._get_instance_orientation(
                orientation, r['instance_segmentation_gt_foreground'], None)
        return r

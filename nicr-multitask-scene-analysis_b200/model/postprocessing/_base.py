# -*- coding: utf-8 -*-
"""Dispatch + full-resolution helper shared by the post-processing classes.

API parity: `PostprocessingBase.postprocess(data, batch, is_training=True)`
(reference: model/postprocessing/base.py:13-40) and
`DensePostprocessingBase._crop_to_valid_region_and_resize_prediction`
(reference: model/postprocessing/dense_base.py:15-58).
"""
import abc
from typing import Any, Dict, Tuple

import torch


class PostprocessingBase(abc.ABC):
    def postprocess(self, data, batch: Dict[str, Any], is_training: bool = True) -> Dict[str, Any]:
        handler = self._postprocess_training if is_training else self._postprocess_inference
        return handler(data, batch)

    @abc.abstractmethod
    def _postprocess_training(self, data, batch):
        ...

    def _postprocess_inference(self, data, batch):
        return self._postprocess_training(data, batch)


class DensePostprocessingBase(PostprocessingBase):
    @staticmethod
    def _is_identity_resize(hw: Tuple[int, int], valid_region_slices, shape) -> bool:
        """True when cropping to the valid region and resizing to `shape` changes nothing
        (network resolution == dataset resolution: every BASELINE configuration)."""
        h, w = hw
        sl_h, sl_w = valid_region_slices
        return (range(h)[sl_h] == range(h) and range(w)[sl_w] == range(w) and
                tuple(shape) == (h, w))

    @staticmethod
    def _crop_geometry(hw: Tuple[int, int], valid_region_slices):
        """(y0, x0, hc, wc) of the valid region inside an (h, w) plane."""
        h, w = hw
        ys, xs = range(h)[valid_region_slices[0]], range(w)[valid_region_slices[1]]
        if ys.step != 1 or xs.step != 1 or len(ys) == 0 or len(xs) == 0:
            raise ValueError('valid region slices must be non-empty with step 1')
        return ys.start, xs.start, len(ys), len(xs)

    def _crop_to_valid_region_and_resize_prediction(
        self,
        prediction: torch.Tensor,
        valid_region_slices: Tuple[slice, slice],
        shape: Tuple[int, int],
        mode: str = 'nearest'
    ) -> torch.Tensor:
        """Crop `...xHxW` to the valid region, then resize to `shape` (h, w): one launch of
        `npb_resize_nearest` / `npb_resize_bilinear` (csrc/resize.cu).  When the shapes
        already agree the cropped view itself is returned, like the reference does."""
        from ctypes import c_int
        from ... import _lib
        sl_h, sl_w = valid_region_slices
        cropped = prediction[..., sl_h, sl_w]
        if tuple(shape) == tuple(cropped.shape[-2:]):
            return cropped
        src = _lib.require_cuda(prediction, 'prediction')
        if src.dtype == torch.bool:
            src = src.view(torch.uint8)
        h_in, w_in = src.shape[-2:]
        y0, x0, hc, wc = self._crop_geometry((h_in, w_in), valid_region_slices)
        planes = src.numel() // (h_in * w_in)
        out = torch.empty(tuple(src.shape[:-2]) + tuple(shape), dtype=src.dtype, device=src.device)
        geom = (c_int(planes), c_int(h_in), c_int(w_in), c_int(y0), c_int(x0), c_int(hc), c_int(wc),
                c_int(shape[0]), c_int(shape[1]))
        if mode == 'nearest':
            _lib.check(_lib.lib().npb_resize_nearest(
                _lib.ptr(src), c_int(src.element_size()), *geom, _lib.ptr(out),
                _lib.stream_ptr(src.device)), 'npb_resize_nearest')
        elif mode == 'bilinear':
            if src.dtype != torch.float32:
                raise TypeError('bilinear resize expects float32')
            _lib.check(_lib.lib().npb_resize_bilinear(
                _lib.ptr(src), *geom, _lib.ptr(out), _lib.stream_ptr(src.device)),
                'npb_resize_bilinear')
        else:
            raise ValueError(f"unsupported resize mode '{mode}'")
        return out.view(torch.bool) if prediction.dtype == torch.bool else out

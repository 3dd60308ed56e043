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
import torch.nn.functional as F


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
        return (tuple(range(h)[sl_h]) == tuple(range(h)) and
                tuple(range(w)[sl_w]) == tuple(range(w)) and tuple(shape) == (h, w))

    def _crop_to_valid_region_and_resize_prediction(
        self,
        prediction: torch.Tensor,
        valid_region_slices: Tuple[slice, slice],
        shape: Tuple[int, int],
        mode: str = 'nearest'
    ) -> torch.Tensor:
        """Crop `...xHxW` to the valid region, then resize to `shape` (h, w).  Identity
        (the cropped view itself) when the shapes already agree.  The general resize is
        SURVEY.md section 8(f) item 1 ("next"): it still runs through torch's interpolate,
        integer maps via an exact f32 round trip like the reference."""
        sl_h, sl_w = valid_region_slices
        cropped = prediction[..., sl_h, sl_w]
        if tuple(shape) == tuple(cropped.shape[-2:]):
            return cropped
        x = cropped.unsqueeze(1) if cropped.ndim == 3 else cropped
        x = x if x.is_floating_point() else x.to(torch.float32)
        extra = {} if mode == 'nearest' else {'align_corners': False}
        x = F.interpolate(x, size=tuple(shape), mode=mode, **extra).to(cropped.dtype)
        return x.squeeze(1) if cropped.ndim == 3 else x

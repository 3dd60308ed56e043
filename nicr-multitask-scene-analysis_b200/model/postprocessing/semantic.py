# -*- coding: utf-8 -*-
"""Semantic post-processing (reference: model/postprocessing/semantic.py:17-82).

The class index map comes from `npb_semantic_argmax` (csrc/semantic.cu): one streaming
pass over the logits producing uint8 classes and, on demand, the soft-max probability of
the winner.  The (B,C,H,W) soft-max tensor and the int64 twin of the class map are part of
the reference's result dict; they are deferred entries here (see _results.ResultDict).
"""
from ctypes import c_int, c_int64

import torch

from ... import _lib
from ..._results import ResultDict
from ...utils.fullres import fullres_key, valid_region_and_fullres_shape
from ._base import DensePostprocessingBase


def semantic_argmax(logits: torch.Tensor, with_score: bool = False):
    """logits (B,C,H,W) f32 CUDA -> (classes uint8 (B,H,W), score f32 (B,H,W) | None)."""
    logits = _lib.require_cuda(logits, 'semantic logits', torch.float32, 4)
    B, C, H, W = logits.shape
    sem = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device)
    score = torch.empty((B, H, W), dtype=torch.float32, device=logits.device) if with_score else None
    _lib.check(_lib.lib().npb_semantic_argmax(
        _lib.ptr(logits), c_int(B), c_int(C), c_int(H), c_int(W), _lib.ptr(sem), _lib.ptr(score),
        _lib.stream_ptr(logits.device)), 'npb_semantic_argmax')
    return sem, score


def softmax_scores(logits: torch.Tensor) -> torch.Tensor:
    """softmax over the class axis of (B,C,H,W) f32 logits (`npb_softmax`)."""
    logits = _lib.require_cuda(logits, 'semantic logits', torch.float32, 4)
    B, C, H, W = logits.shape
    probs = torch.empty_like(logits)
    _lib.check(_lib.lib().npb_softmax(_lib.ptr(logits), c_int(B), c_int(C), c_int(H), c_int(W),
                                      _lib.ptr(probs), _lib.stream_ptr(logits.device)),
               'npb_softmax')
    return probs


def semantic_argmax_resized(logits: torch.Tensor, crop_geometry, shape):
    """Bilinear resize (align_corners=False) of the cropped logits to `shape` fused with the
    arg-max: -> (classes uint8 (B,h,w), score f32 (B,h,w)); `npb_semantic_argmax_resized`."""
    logits = _lib.require_cuda(logits, 'semantic logits', torch.float32, 4)
    B, C, H, W = logits.shape
    y0, x0, hc, wc = crop_geometry
    h, w = shape
    sem = torch.empty((B, h, w), dtype=torch.uint8, device=logits.device)
    score = torch.empty((B, h, w), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.lib().npb_semantic_argmax_resized(
        _lib.ptr(logits), c_int(B), c_int(C), c_int(H), c_int(W), c_int(y0), c_int(x0), c_int(hc),
        c_int(wc), c_int(h), c_int(w), _lib.ptr(sem), _lib.ptr(score),
        _lib.stream_ptr(logits.device)), 'npb_semantic_argmax_resized')
    return sem, score


def widen_u8(x: torch.Tensor, add: int = 0) -> torch.Tensor:
    """uint8 map -> int64 map (+ add) on the device (`npb_widen_u8`)."""
    x = _lib.require_cuda(x, 'uint8 map', torch.uint8)
    out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    _lib.check(_lib.lib().npb_widen_u8(_lib.ptr(x), c_int64(x.numel()), c_int64(add),
                                       _lib.ptr(out), _lib.stream_ptr(x.device)), 'npb_widen_u8')
    return out


class SemanticPostprocessing(DensePostprocessingBase):
    def __init__(self, **kwargs):
        super().__init__()

    def _postprocess_training(self, data, batch):
        output, side_outputs = data
        return {'semantic_output': output, 'semantic_side_outputs': side_outputs}

    def _fill_inference_entries(self, r: ResultDict, logits: torch.Tensor, batch,
                                sem_u8: torch.Tensor = None) -> ResultDict:
        """Adds the inference entries of semantic.py:52-80 for `logits`; `sem_u8` may be
        supplied by a caller that already computed the class map (fused panoptic path)."""
        cache = {'sem': sem_u8, 'score': None}

        def classes():
            if cache['sem'] is None:
                cache['sem'], _ = semantic_argmax(logits)
            return cache['sem']

        def score():
            if cache['score'] is None:
                sem, cache['score'] = semantic_argmax(logits, with_score=True)
                if cache['sem'] is None:
                    cache['sem'] = sem
            return cache['score']

        r['_semantic_segmentation_idx_u8'] = classes() if sem_u8 is None else sem_u8
        r.defer('semantic_softmax_scores', lambda: softmax_scores(logits))
        r.defer('semantic_segmentation_score', score)
        r.defer('semantic_segmentation_idx', lambda: widen_u8(classes()))

        crop, shape = valid_region_and_fullres_shape(batch, 'semantic')
        if self._is_identity_resize(tuple(logits.shape[-2:]), crop, shape):
            r['semantic_output_fullres'] = logits
            for k in ('semantic_softmax_scores', 'semantic_segmentation_score',
                      'semantic_segmentation_idx'):
                r.alias(fullres_key(k), k)
        else:
            # network resolution != dataset resolution (semantic.py:63-72): the class map and
            # its score come from ONE fused kernel (bilinear resize on the fly + arg-max); the
            # resized logits / soft-max tensors are only materialised if somebody reads them
            fcache = {}
            geom = self._crop_geometry(tuple(logits.shape[-2:]), crop)

            def full_pair():
                if 'sem' not in fcache:
                    fcache['sem'], fcache['score'] = semantic_argmax_resized(logits, geom, shape)
                return fcache

            def full_logits():
                if 'logits' not in fcache:
                    fcache['logits'] = self._crop_to_valid_region_and_resize_prediction(
                        logits, crop, shape, mode='bilinear')
                return fcache['logits']

            r.defer('semantic_output_fullres', full_logits)
            r.defer('semantic_softmax_scores_fullres', lambda: softmax_scores(full_logits()))
            r.defer('semantic_segmentation_score_fullres', lambda: full_pair()['score'])
            r.defer('semantic_segmentation_idx_fullres', lambda: widen_u8(full_pair()['sem']))
        return r

    def _postprocess_inference(self, data, batch):
        output, side_outputs = data
        r = ResultDict(semantic_output=output, semantic_side_outputs=side_outputs)
        return self._fill_inference_entries(r, output, batch)

# -*- coding: utf-8 -*-
"""Mean intersection over union from an int64 confusion matrix kept on the GPU
(API of metric/miou.py:9-94).  `update` is one launch of `npb_confmat_update`
(csrc/eval.cu: 128-bit streaming loads, warp-aggregated, shared-memory privatised histogram)."""
from ctypes import c_int, c_int64

import torch

from .. import _lib
from ._state import MetricState


class MeanIntersectionOverUnion(MetricState):
    def __init__(self, n_classes: int, ignore_first_class: bool = False, device=None) -> None:
        super().__init__(device)
        self._n_classes = n_classes
        self._ignore_first_class = ignore_first_class
        self.add_state('confmat', torch.zeros((n_classes, n_classes), dtype=torch.int64),
                       dist_reduce_fx='sum')
        self._status = None

    def _status_word(self) -> torch.Tensor:
        if self._status is None or self._status.device != self.confmat.device:
            self._status = torch.zeros(1, dtype=torch.int32, device=self.confmat.device)
        return self._status

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """confmat[target, pred] += 1 for every element (rows = target, miou.py:50-56).
        Any integer dtype; values must lie in [0, n_classes)."""
        self._update(preds, target, 'npb_confmat_update')

    def update_nonvoid(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """`update(preds[target != 0], target[target != 0] - 1)` without the masked copies:
        the call of SemanticTaskHelper.validation_step (task_helper/semantic.py:126-131).
        `preds` in [0, n_classes) at the non-void elements, `target` in [0, n_classes]."""
        self._update(preds, target, 'npb_confmat_update_nonvoid')

    def _update(self, preds: torch.Tensor, target: torch.Tensor, entry: str) -> None:
        if not self.confmat.is_cuda:
            raise RuntimeError('MeanIntersectionOverUnion.update needs its state on a CUDA device')
        dev = self.confmat.device
        preds = _lib.require_cuda(preds.to(dev), 'preds')
        target = _lib.require_cuda(target.to(dev), 'target')
        if preds.numel() != target.numel():
            raise ValueError('preds and target differ in size')
        status = self._status_word()
        _lib.check(getattr(_lib.lib(), entry)(
            _lib.ptr(preds), c_int(_lib.dtype_code(preds)), _lib.ptr(target),
            c_int(_lib.dtype_code(target)), c_int64(preds.numel()), c_int(self._n_classes),
            _lib.ptr(self.confmat), _lib.ptr(status), _lib.stream_ptr(dev)), entry)

    def check_status(self) -> None:
        if self._status is not None:
            code = int(self._status.item())
            self._status.zero_()
            _lib.raise_for_status([code], 'MeanIntersectionOverUnion.update')

    def reset(self) -> None:
        super().reset()
        if self._status is not None:
            self._status.zero_()

    def compute(self, return_ious: bool = False):
        """miou.py:58-94 on the (rank-summed) confusion matrix; float32 like the reference.
        Results are CPU tensors."""
        self.check_status()
        cm = self.host_states()['confmat']
        tp = torch.diag(cm).float()
        sum_pred = cm.sum(dim=0).float()
        sum_gt = cm.sum(dim=1).float()
        if self._ignore_first_class:
            # void row / column dropped; void-GT pixels do not count as predictions either
            tp, sum_pred, sum_gt = tp[1:], sum_pred[1:] - cm[0, 1:].float(), sum_gt[1:]
        has_gt = sum_gt != 0
        iou = tp[has_gt] / (sum_pred[has_gt] + sum_gt[has_gt] - tp[has_gt])
        miou = torch.mean(iou)
        if not return_ious:
            return miou
        ious = torch.full((self._n_classes,), torch.nan, dtype=torch.float32)
        where = has_gt.nonzero(as_tuple=True)[0] + (1 if self._ignore_first_class else 0)
        ious[where] = iou
        return miou, ious

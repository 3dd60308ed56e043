# -*- coding: utf-8 -*-
"""`SemanticTaskHelper`: validation of the semantic segmentation
(task_helper/semantic.py:21-161 without losses and visualisation examples).

The reference's step builds `mask = target != 0`, the masked copies `preds[mask]` and
`target[mask] - 1`, moves both to the CPU and calls `MeanIntersectionOverUnion.update`
(task_helper/semantic.py:126-131).  Here the full-resolution maps stay on the device and ONE
launch of `npb_confmat_update_nonvoid` skips the void elements while it counts; state,
result keys and the (artifacts, examples, logs) triple are the reference's.
"""
from typing import Any, Dict, Optional, Tuple

import torch

from ..metric import MeanIntersectionOverUnion
from ..utils.fullres import fullres_key
from .base import TaskHelperBase, get_fullres


class SemanticTaskHelper(TaskHelperBase):
    def __init__(self, n_classes: int, class_weights: Optional[Any] = None,
                 label_smoothing: float = 0.0, disable_multiscale_supervision: bool = False,
                 examples_cmap: Optional[Any] = None) -> None:
        """Signature of task_helper/semantic.py:22-29; `n_classes` is WITHOUT void.  The loss
        and colour-map arguments are accepted and ignored (training and visualisation are
        outside this package)."""
        super().__init__()
        self._n_classes = int(n_classes)

    def initialize(self, device: torch.device) -> None:
        super().initialize(device)
        self._metric_iou = MeanIntersectionOverUnion(n_classes=self._n_classes,
                                                     device=self.device)
        self._metric_iou.reset()

    def validation_step(self, batch: Dict[str, Any], batch_idx: int,
                        predictions_post: Dict[str, Any]) -> Tuple[Dict, Dict]:
        return self._timed('semantic_step_time', self._validation_step, batch, batch_idx,
                           predictions_post)

    def _validation_step(self, batch, batch_idx, predictions_post):
        target = self._dev(get_fullres(batch, 'semantic'))            # 0 = void
        preds = self._dev(predictions_post[fullres_key('semantic_segmentation_idx')])
        self._metric_iou.update_nonvoid(preds=preds, target=target)
        return {}, {}

    def validation_epoch_end(self):
        return self._timed('semantic_epoch_end_time', self._validation_epoch_end)

    def _validation_epoch_end(self):
        miou, ious = self._metric_iou.compute(return_ious=True)
        logs = {'semantic_miou': miou}
        artifacts = {'semantic_cm': self._metric_iou.confmat.clone(),
                     'semantic_ious_per_class': ious.clone()}
        self._metric_iou.reset()
        assert self._metric_iou.confmat.sum() == 0
        return artifacts, self._examples, logs

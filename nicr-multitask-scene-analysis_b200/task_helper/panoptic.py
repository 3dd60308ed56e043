# -*- coding: utf-8 -*-
"""`PanopticTaskHelper`: validation of the merged panoptic prediction
(task_helper/panoptic.py:28-212 without the visualisation examples).

The reference's step moves prediction and target to the CPU, runs
`PanopticQualityWithOrientationMAE.update` (process pool) and then
`MeanIntersectionOverUnion.update(pred // max_instances, semantic_target)`.  Here both metric
states are fed by ONE launch of `npb_pq_update` over device-resident maps
(metric/fused.py); states, result keys and the (artifacts, examples, logs) triple are the
reference's.
"""
from typing import Any, Dict, Optional, Sequence, Tuple

import torch

from ..metric import (MeanIntersectionOverUnion, PanopticEvaluation,
                      PanopticQualityWithOrientationMAE)
from ..utils.fullres import fullres_key
from .base import TaskHelperBase, get_fullres


class PanopticTaskHelper(TaskHelperBase):
    def __init__(self, semantic_n_classes: int, semantic_classes_is_thing: Sequence[bool],
                 semantic_label_list: Optional[Any] = None) -> None:
        """`semantic_n_classes` and `semantic_classes_is_thing` include void
        (task_helper/panoptic.py:29-40); `semantic_label_list` only feeds the reference's
        colour generator and is kept for signature compatibility."""
        super().__init__()
        self._semantic_n_classes = int(semantic_n_classes)
        self._semantic_classes_is_thing = tuple(bool(t) for t in semantic_classes_is_thing)
        self._semantic_label_list = semantic_label_list
        self._max_instances_per_category = 1 << 16          # task_helper/panoptic.py:42
        self._thing_ids = [i for i, t in enumerate(self._semantic_classes_is_thing) if t]
        self._with_orientation = False

    def initialize(self, device: torch.device) -> None:
        super().initialize(device)
        self._mae_pq_deeplab = PanopticQualityWithOrientationMAE(
            num_categories=self._semantic_n_classes, ignored_label=0,
            max_instances_per_category=self._max_instances_per_category, offset=256 ** 3,
            is_thing=self._semantic_classes_is_thing, device=self.device)
        self._metric_iou = MeanIntersectionOverUnion(
            n_classes=self._semantic_n_classes, ignore_first_class=True, device=self.device)
        self._metric_iou.reset()
        self._evaluation = PanopticEvaluation(self._mae_pq_deeplab, self._metric_iou)

    @property
    def evaluation(self) -> PanopticEvaluation:
        """The PQ + mIoU pair of this helper, e.g. for `PanopticPostprocessing.fuse_evaluation`."""
        return self._evaluation

    def validation_step(self, batch: Dict[str, Any], batch_idx: int,
                        predictions_post: Dict[str, Any]) -> Tuple[Dict, Dict]:
        return self._timed('panoptic_step_time', self._validation_step, batch, batch_idx,
                           predictions_post)

    def _validation_step(self, batch, batch_idx, predictions_post):
        self._with_orientation = 'orientations_present' in batch
        if self._with_orientation:
            orientations_results = predictions_post['orientations_panoptic_segmentation_deeplab_instance']
            orientations_targets = batch['orientations_present']
        else:
            orientations_results = orientations_targets = None

        if predictions_post.get('_panoptic_evaluation_fused'):
            # the post-processing kernel already fed PQ + mIoU (PanopticPostprocessing.
            # fuse_evaluation); only the MAAE loop over the matched instances is left
            matches = predictions_post.get('_panoptic_matches')
            if self._with_orientation and matches is not None:
                self._evaluation.update_mae_from_matches(
                    matches, orientations_results,
                    predictions_post['panoptic_segmentation_deeplab_ids'], orientations_targets,
                    batch.get('panoptic_ids_to_instance_dict'))
            return {}, {}

        panoptic_targets = self._dev(get_fullres(batch, 'panoptic'))
        panoptic_preds = predictions_post[fullres_key('panoptic_segmentation_deeplab')]
        semantic_targets = self._dev(get_fullres(batch, 'semantic'))
        # PQ (+ matches for the MAAE loop) and the confusion matrix of `pred // L` in one pass
        self._evaluation.update_with_orientation(
            panoptic_preds=self._dev(panoptic_preds),
            orientation_preds=orientations_results,
            panoptic_preds_id_dicts=predictions_post['panoptic_segmentation_deeplab_ids'],
            panoptic_target=panoptic_targets,
            orientation_target=orientations_targets,
            panoptic_target_id_dicts=batch.get('panoptic_ids_to_instance_dict'),
            semantic_target=semantic_targets)
        return {}, {}

    def validation_epoch_end(self):
        return self._timed('panoptic_epoch_end_time', self._validation_epoch_end)

    def _validation_epoch_end(self):
        artifacts: Dict[str, Any] = {}
        logs: Dict[str, Any] = {}
        self._split_results('panoptic', self._mae_pq_deeplab.compute(suffix='_deeplab'),
                            artifacts, logs)
        self._mae_pq_deeplab.reset()

        artifacts['panoptic_deeplab_semantic_cm'] = self._metric_iou.confmat.clone()
        miou, ious = self._metric_iou.compute(return_ious=True)
        logs['panoptic_deeplab_semantic_miou'] = miou
        artifacts['panoptic_deeplab_semantic_ious_per_class'] = ious
        self._metric_iou.reset()
        assert self._metric_iou.confmat.sum() == 0
        return artifacts, self._examples, logs

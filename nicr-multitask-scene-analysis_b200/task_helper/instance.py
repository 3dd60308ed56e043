# -*- coding: utf-8 -*-
"""`InstanceTaskHelper`: stand-alone evaluation of the instance segmentation
(task_helper/instance.py:289-436 without losses and visualisation examples).

The predicted instances (grouped inside the ground-truth foreground) are merged with the
ground-truth semantic map by `deeplab_merge_batch` and scored with PQ against the panoptic
target, so the number tells how good the instances are under a perfect semantic
segmentation.  Merge and PQ run on the device; the reference moves everything to the CPU.
"""
from typing import Any, Dict, Sequence, Tuple

import torch

from ..metric import MeanAbsoluteAngularError, PanopticQualityWithOrientationMAE
from ..utils.fullres import fullres_key
from ..utils.panoptic_merge import deeplab_merge_batch
from .base import TaskHelperBase, get_fullres


class InstanceTaskHelper(TaskHelperBase):
    def __init__(self, semantic_n_classes: int, semantic_classes_is_thing: Sequence[bool],
                 loss_name_instance_center: str = 'mse',
                 disable_multiscale_supervision: bool = False) -> None:
        """Signature of task_helper/instance.py:36-42; the two loss arguments are accepted
        and ignored (training is outside this package)."""
        super().__init__()
        self._semantic_n_classes = int(semantic_n_classes)
        self._semantic_classes_is_thing = tuple(bool(t) for t in semantic_classes_is_thing)
        self._max_instances_per_category = 1 << 16          # task_helper/instance.py:58
        self._thing_ids = [i for i, t in enumerate(self._semantic_classes_is_thing) if t]
        self._with_orientation = False

    def initialize(self, device: torch.device) -> None:
        super().initialize(device)
        self._mae_pq_deeplab = PanopticQualityWithOrientationMAE(
            num_categories=self._semantic_n_classes, ignored_label=0,
            max_instances_per_category=self._max_instances_per_category, offset=256 ** 3,
            is_thing=self._semantic_classes_is_thing, device=self.device)
        self._mae_gt = MeanAbsoluteAngularError(device=self.device)

    def validation_step(self, batch: Dict[str, Any], batch_idx: int,
                        predictions_post: Dict[str, Any]) -> Tuple[Dict, Dict]:
        return self._timed('instance_step_time', self._validation_step, batch, batch_idx,
                           predictions_post)

    def _validation_step(self, batch, batch_idx, predictions_post):
        self._with_orientation = 'orientations_present' in batch
        if self._with_orientation:
            orientations_results = \
                predictions_post['orientations_instance_segmentation_gt_orientation_foreground']
            orientations_full_gt = \
                predictions_post['orientations_gt_instance_gt_orientation_foreground']
            orientations_targets = batch['orientations_present']
            self._mae_gt.update(orientations_full_gt, orientations_targets)
        else:
            orientations_results = orientations_targets = None

        semantic_batch = self._dev(get_fullres(batch, 'semantic'))
        instance_batch = self._dev(get_fullres(batch, 'instance'))
        instance_result = self._dev(
            predictions_post[fullres_key('instance_segmentation_gt_foreground')])
        instance_foreground = instance_batch != 0
        panoptic_targets = self._dev(get_fullres(batch, 'panoptic'))

        # ground-truth semantic + predicted instances -> panoptic prediction
        panoptic_preds, panoptic_id_dicts = deeplab_merge_batch(
            semantic_batch, instance_result, instance_foreground,
            self._max_instances_per_category, self._thing_ids, 0)
        self._mae_pq_deeplab.update(panoptic_preds, orientations_results, panoptic_id_dicts,
                                    panoptic_targets, orientations_targets,
                                    batch.get('panoptic_ids_to_instance_dict'))
        return {}, {}

    def validation_epoch_end(self):
        return self._timed('instance_epoch_end_time', self._validation_epoch_end)

    def _validation_epoch_end(self):
        artifacts: Dict[str, Any] = {}
        logs: Dict[str, Any] = {}
        self._split_results('instance', self._mae_pq_deeplab.compute(suffix='_deeplab'),
                            artifacts, logs)
        self._mae_pq_deeplab.reset()
        if self._with_orientation:
            mae_gt_rad, mae_gt_deg = self._mae_gt.compute()
            logs['orientation_mae_gt_rad'] = mae_gt_rad
            logs['orientation_mae_gt_deg'] = mae_gt_deg
            self._mae_gt.reset()
        return artifacts, self._examples, logs

# -*- coding: utf-8 -*-
"""One-pass evaluation: PQ and the mIoU confusion matrix from a single read of the
predicted panoptic map.

The reference's validation step (task_helper/panoptic.py:104-126) calls
`PanopticQualityWithOrientationMAE.update(pred, target)` and then
`MeanIntersectionOverUnion.update(pred // max_instances, semantic_target)`: the prediction
is read twice and `pred // L` is materialised as an int64 tensor in between.
`PanopticEvaluation.update` feeds both metric objects from one launch of `npb_pq_update`
(17 bytes per pixel: int64 prediction, int64 panoptic target, uint8 semantic target); the
states it leaves in the two metric objects are identical to those of the two separate calls.
"""

import torch

from .miou import MeanIntersectionOverUnion
from .pq import PanopticQuality


class PanopticEvaluation:
    def __init__(self, pq: PanopticQuality, miou: MeanIntersectionOverUnion):
        assert miou._n_classes <= 256
        self.pq = pq
        self.miou = miou

    def update(self, panoptic_preds: torch.Tensor, panoptic_targets: torch.Tensor,
               semantic_targets: torch.Tensor, want_matches: bool = False):
        """(B,H,W) int64, (B,H,W) int64, (B,H,W) uint8.  Returns (matches, n_matches) device
        tensors when `want_matches` (for the MAAE loop), else None."""
        if panoptic_preds.shape[0] == 0:        # an empty batch adds nothing
            return (None, None) if want_matches else None
        if semantic_targets.dtype != torch.uint8:
            semantic_targets = semantic_targets.to(torch.uint8)
        matches, n_matches, _ = self.pq._launch(
            panoptic_preds, panoptic_targets, sem_target=semantic_targets,
            confmat=self.miou.confmat, want_matches=want_matches)
        return (matches, n_matches) if want_matches else None

    def eval_args(self, panoptic_targets: torch.Tensor, semantic_targets: torch.Tensor,
                  want_matches: bool = False):
        """Evaluation half of a fused post-processing + evaluation call
        (`PanopticPostprocessing.fuse_evaluation`): returns (`_lib.EvalArgs`, tensors to keep
        alive until the call has been issued; 'matches' / 'n_matches' when requested)."""
        if semantic_targets.dtype != torch.uint8:
            semantic_targets = semantic_targets.to(torch.uint8)
        return self.pq._eval_args(panoptic_targets, sem_target=semantic_targets,
                                  confmat=self.miou.confmat, want_matches=want_matches)

    def update_with_orientation(self, panoptic_preds: torch.Tensor, orientation_preds,
                                panoptic_preds_id_dicts, panoptic_target: torch.Tensor,
                                orientation_target, panoptic_target_id_dicts,
                                semantic_target: torch.Tensor) -> None:
        """The validation step of the reference's panoptic task helper
        (task_helper/panoptic.py:104-126) in one call: the arguments of
        `PanopticQualityWithOrientationMAE.update` (mae.py:84-127) plus the semantic target
        of the mIoU update.  `self.pq` must carry the MAAE state."""
        assert panoptic_preds.ndim == 3
        assert len(panoptic_target) == len(panoptic_preds)
        with_mae = orientation_preds is not None and orientation_target is not None
        out = self.update(panoptic_preds, panoptic_target, semantic_target, want_matches=with_mae)
        if not with_mae:
            return
        if out[0] is None:
            return
        self.update_mae_from_matches(out, orientation_preds, panoptic_preds_id_dicts,
                                     orientation_target, panoptic_target_id_dicts)

    def update_mae_from_matches(self, matches, orientation_preds, panoptic_preds_id_dicts,
                                orientation_target, panoptic_target_id_dicts) -> None:
        """MAAE part of an update whose PQ part already ran: `matches` = (pairs, counts) device
        tensors of that launch (mae.py:113-127)."""
        pairs_d, n_matches = matches
        self.pq.update_mae_batch(pairs_d, n_matches, orientation_preds, panoptic_preds_id_dicts,
                                 orientation_target, panoptic_target_id_dicts)

    def reset(self) -> None:
        self.pq.reset()
        self.miou.reset()

    def compute(self, suffix: str = ''):
        out = self.pq.compute(suffix=suffix)
        miou, ious = self.miou.compute(return_ious=True)
        out[f'semantic{suffix}_miou'] = miou
        out[f'semantic{suffix}_ious_per_class'] = ious
        return out

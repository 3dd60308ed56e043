# -*- coding: utf-8 -*-
"""Mean absolute angular error and PQ + MAAE over matched instances (API of
metric/mae.py:16-172).  The PQ part runs in `npb_pq_update`; the MAAE part works on the
matched (gt id, pred id) pairs the kernel returns and on the python id / orientation dicts of
the reference's interface (O(#instances) look-ups on the host), but evaluates the errors of a
whole batch with one float32 vector operation and adds them to the device state once --
the reference's loop costs six scalar tensor operations and a state update per matched pair."""
from typing import Dict, List, Tuple

import torch

from .pq import PanopticQuality
from ._state import MetricState


def abs_angle_error_rad(pred_angle: torch.Tensor, target_angle: torch.Tensor) -> torch.Tensor:
    """Smallest absolute difference between two angles (mae.py:16-30), in [0, pi]."""
    two_pi = 2 * torch.pi
    diff = pred_angle % two_pi - target_angle % two_pi
    return torch.abs((diff + torch.pi) % two_pi - torch.pi)


class _AngularErrorMixin:
    def _add_angular_state(self):
        self.add_state('sum_angular_error', torch.tensor(0, dtype=torch.float64),
                       dist_reduce_fx='sum')
        self.add_state('n_elements', torch.tensor(0, dtype=torch.int64), dist_reduce_fx='sum')

    def _add_errors(self, pred_angles: List[float], target_angles: List[float]) -> None:
        """Errors of a whole batch of (prediction, target) pairs with ONE vector operation and
        ONE update of the device state (the reference issues six scalar tensor operations
        and a state update per pair, mae.py:55-58, 157-162).  Element-wise float32 arithmetic
        like `torch.tensor(python float)` there; the float64 sum runs over the pairs in order."""
        if not pred_angles:
            return
        err = abs_angle_error_rad(torch.tensor(pred_angles, dtype=torch.float32),
                                  torch.tensor(target_angles, dtype=torch.float32))
        total = 0.0
        for e in err.tolist():          # python floats: sequential float64 adds
            total += e
        self.sum_angular_error += total
        self.n_elements += len(pred_angles)

    def _mean_error(self, states) -> Tuple[torch.Tensor, torch.Tensor]:
        rad = states['sum_angular_error'].cpu() / states['n_elements'].cpu()
        return rad, torch.rad2deg(rad)


class MeanAbsoluteAngularError(MetricState, _AngularErrorMixin):
    def __init__(self, device=None, **kwargs):
        super().__init__(device)
        self._add_angular_state()

    def update(self, orientation_preds: List[Dict], orientation_target: List[Dict]) -> None:
        pred_angles, target_angles = [], []
        for preds, targets in zip(orientation_preds, orientation_target):
            for key, pred_angle in preds.items():
                pred_angles.append(pred_angle)
                target_angles.append(targets[key])
        self._add_errors(pred_angles, target_angles)

    def compute(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._mean_error(self.synced_states())


class PanopticQualityWithOrientationMAE(PanopticQuality, _AngularErrorMixin):
    """PQ plus the mean absolute angular error over matched instances."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._add_angular_state()

    def update(self, panoptic_preds: torch.Tensor, orientation_preds, panoptic_preds_id_dicts,
               panoptic_target: torch.Tensor, orientation_target, panoptic_target_id_dicts):
        assert panoptic_preds.ndim == 3
        assert len(panoptic_target) == len(panoptic_preds)
        with_mae = orientation_preds is not None and orientation_target is not None
        if panoptic_preds.shape[0] == 0:        # an empty batch adds nothing
            return
        matches, n_matches, _ = self._launch(panoptic_preds, panoptic_target,
                                             want_matches=with_mae)
        if not with_mae:
            return
        self.update_mae_batch(matches, n_matches, orientation_preds, panoptic_preds_id_dicts,
                              orientation_target, panoptic_target_id_dicts)

    def update_mae_batch(self, matches: torch.Tensor, n_matches: torch.Tensor, orientation_preds,
                         panoptic_preds_id_dicts, orientation_target,
                         panoptic_target_id_dicts) -> None:
        """MAAE part of a batch whose PQ part has been launched: `matches` (B, cap, 2) /
        `n_matches` (B) are the device outputs of that launch (mae.py:113-127).  One read-back
        of the matched pairs, dict look-ups on the host, one vector operation for all errors."""
        self.check_status()
        counts = n_matches.cpu().tolist()
        pairs = matches.cpu()
        pred_angles, target_angles = [], []
        for b, n in enumerate(counts):
            self._collect_mae_pairs(orientation_preds[b], panoptic_preds_id_dicts[b],
                                    orientation_target[b], panoptic_target_id_dicts[b],
                                    pairs[b, :n].tolist(), pred_angles, target_angles)
        self._add_errors(pred_angles, target_angles)

    @staticmethod
    def _collect_mae_pairs(orientation_preds, panoptic_preds_id_dicts, orientation_target,
                           panoptic_target_id_dicts, matching, pred_angles, target_angles) -> None:
        """mae.py:129-162: a matched pair contributes when both sides carry an orientation."""
        for target_id, pred_id in matching:
            if target_id == 0:
                continue            # stuff / void / background
            target_instance = panoptic_target_id_dicts.get(target_id)
            pred_instance = panoptic_preds_id_dicts.get(pred_id)
            if target_instance is None or target_instance not in orientation_target:
                continue
            if pred_instance is None or pred_instance not in orientation_preds:
                continue
            pred_angles.append(orientation_preds[pred_instance])
            target_angles.append(orientation_target[target_instance])

    def update_mae(self, orientation_preds, panoptic_preds_id_dicts, orientation_target,
                   panoptic_target_id_dicts, matching):
        """The reference's per-frame entry (mae.py:129-162)."""
        pred_angles, target_angles = [], []
        self._collect_mae_pairs(orientation_preds, panoptic_preds_id_dicts, orientation_target,
                                panoptic_target_id_dicts, matching, pred_angles, target_angles)
        self._add_errors(pred_angles, target_angles)

    def compute(self, suffix: str = '') -> Dict:
        r = super().compute(suffix=suffix)
        rad, deg = self._mean_error(self.synced_states())
        r[f'mae{suffix}_rad'] = rad
        r[f'mae{suffix}_deg'] = deg
        return r

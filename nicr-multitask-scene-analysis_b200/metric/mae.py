# -*- coding: utf-8 -*-
"""Mean absolute angular error and PQ + MAAE over matched instances (API of
metric/mae.py:16-172).  The PQ part runs in `npb_pq_update`; the MAAE loop works on the
matched (gt id, pred id) pairs the kernel returns and on python dicts, exactly like the
reference (O(#instances) host work)."""
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

    def _add_error(self, pred_angle: float, target_angle: float) -> None:
        # float32 scalars like torch.tensor(python float) in the reference (mae.py:55-58)
        err = abs_angle_error_rad(torch.tensor(pred_angle), torch.tensor(target_angle))
        self.sum_angular_error += err.to(self.sum_angular_error.device)
        self.n_elements += 1

    def _mean_error(self, states) -> Tuple[torch.Tensor, torch.Tensor]:
        rad = states['sum_angular_error'].cpu() / states['n_elements'].cpu()
        return rad, torch.rad2deg(rad)


class MeanAbsoluteAngularError(MetricState, _AngularErrorMixin):
    def __init__(self, device=None, **kwargs):
        super().__init__(device)
        self._add_angular_state()

    def update(self, orientation_preds: List[Dict], orientation_target: List[Dict]) -> None:
        for preds, targets in zip(orientation_preds, orientation_target):
            for key, pred_angle in preds.items():
                self._add_error(pred_angle, targets[key])

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
        matches, n_matches, _ = self._launch(panoptic_preds, panoptic_target,
                                             want_matches=with_mae)
        if not with_mae:
            return
        self.check_status()
        counts = n_matches.cpu().tolist()
        pairs = matches.cpu()
        for b, n in enumerate(counts):
            self.update_mae(orientation_preds[b], panoptic_preds_id_dicts[b],
                            orientation_target[b], panoptic_target_id_dicts[b],
                            [tuple(p) for p in pairs[b, :n].tolist()])

    def update_mae(self, orientation_preds, panoptic_preds_id_dicts, orientation_target,
                   panoptic_target_id_dicts, matching):
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
            self._add_error(orientation_preds[pred_instance], orientation_target[target_instance])

    def compute(self, suffix: str = '') -> Dict:
        r = super().compute(suffix=suffix)
        rad, deg = self._mean_error(self.synced_states())
        r[f'mae{suffix}_rad'] = rad
        r[f'mae{suffix}_deg'] = deg
        return r

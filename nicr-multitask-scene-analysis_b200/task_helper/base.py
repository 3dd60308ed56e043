# -*- coding: utf-8 -*-
"""Common part of the evaluation task helpers (interface of task_helper/base.py:79-140
restricted to what the validation path uses)."""
from time import perf_counter
from typing import Any, Dict, Optional, Tuple

import torch

from ..utils.fullres import fullres_key


def get_fullres(batch: Dict[str, Any], key: str) -> Any:
    """`<key>_fullres` entry of the batch or None (data/preprocessing/resize.py:26-27)."""
    return batch.get(fullres_key(key), None)


class TaskHelperBase:
    def __init__(self) -> None:
        self._device: Optional[torch.device] = None
        self._examples: Dict[str, Any] = {}

    @property
    def device(self) -> torch.device:
        if self._device is None:
            raise RuntimeError('task helper used before initialize(device)')
        return self._device

    def initialize(self, device: torch.device) -> None:
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError(f'{type(self).__name__}: metrics run on a CUDA device (there is no '
                               f'CPU path); got {device}')
        self._device = device

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        """Batch entries arrive from the data loader on the host: move them next to the
        predictions (a no-op for tensors already resident)."""
        return t.to(self.device, non_blocking=True)

    def training_step(self, batch, batch_idx: int, predictions_post) -> Tuple[Dict, Dict]:
        return {}, {}

    @staticmethod
    def _timed(key: str, fn, *args, **kwargs):
        """The reference wraps every step in `append_profile_to_logs` (base.py:47-66): the last
        element of the result is the logs dict and receives the host time of the step."""
        start = perf_counter()
        results = fn(*args, **kwargs)
        results[-1][key] = perf_counter() - start
        return results

    @staticmethod
    def _split_results(prefix: str, results: Dict[str, Any], artifacts: Dict, logs: Dict) -> None:
        """scalars -> logs, per-class vectors -> artifacts (task_helper/panoptic.py:193-197)."""
        for key, value in results.items():
            value = torch.as_tensor(value)
            if value.numel() == 1:
                logs[f'{prefix}_{key}'] = value
            else:
                artifacts[f'{prefix}_{key}'] = value

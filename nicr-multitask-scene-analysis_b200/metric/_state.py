# -*- coding: utf-8 -*-
"""Minimal metric-state bookkeeping (what the reference takes from torchmetrics.Metric:
`add_state(..., dist_reduce_fx='sum')`, `reset()`, `.to(device)`), plus the one data-path
collective of this package: a SUM all-reduce of the states over `torch.distributed`
(NCCL on the GPUs, gloo in the CPU tests) when `compute()` is called."""
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


class MetricState:
    """Named tensors with defaults; every state is summed across ranks."""

    def __init__(self, device: Optional[torch.device] = None):
        self._device = torch.device(device) if device is not None else default_device()
        self._defaults: Dict[str, torch.Tensor] = {}

    # ---- torchmetrics-like surface ---------------------------------------------------------
    def add_state(self, name: str, default: torch.Tensor, dist_reduce_fx: str = 'sum') -> None:
        assert dist_reduce_fx == 'sum'
        self._defaults[name] = default.detach().clone().cpu()
        setattr(self, name, default.detach().clone().to(self._device))

    def reset(self) -> None:
        """Back to the defaults IN PLACE: the state tensors keep their addresses, so kernels
        captured in a CUDA graph keep accumulating into the live states."""
        for name, default in self._defaults.items():
            getattr(self, name).copy_(default)

    def to(self, device) -> 'MetricState':
        self._device = torch.device(device)
        for name in self._defaults:
            setattr(self, name, getattr(self, name).to(self._device))
        return self

    @property
    def device(self) -> torch.device:
        return self._device

    def state_names(self) -> List[str]:
        return list(self._defaults)

    # ---- cross-rank reduction --------------------------------------------------------------
    def synced_states(self) -> Dict[str, torch.Tensor]:
        """States summed over all ranks (copies; the local states are left untouched, like
        torchmetrics' sync/unsync around compute()).  Same-dtype states are packed into one
        buffer so a compute() costs one all-reduce per dtype (int64 / float64)."""
        local = {n: getattr(self, n) for n in self._defaults}
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return {n: t.clone() for n, t in local.items()}
        out = {}
        by_dtype: Dict[torch.dtype, List[str]] = {}
        for n, t in local.items():
            by_dtype.setdefault(t.dtype, []).append(n)
        for dtype, names in by_dtype.items():
            flat = torch.cat([local[n].reshape(-1) for n in names])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            off = 0
            for n in names:
                k = local[n].numel()
                out[n] = flat[off:off + k].reshape(local[n].shape).clone()
                off += k
        return out

    def host_states(self) -> Dict[str, torch.Tensor]:
        """`synced_states()` on the host with ONE device->host copy per dtype (every copy is a
        synchronisation: four PQ vectors one by one cost more than the kernels of a small batch)."""
        local = {n: getattr(self, n) for n in self._defaults}
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        out = {}
        by_dtype: Dict[torch.dtype, List[str]] = {}
        for n, t in local.items():
            by_dtype.setdefault(t.dtype, []).append(n)
        for dtype, names in by_dtype.items():
            flat = torch.cat([local[n].reshape(-1) for n in names]) if len(names) > 1 or distributed \
                else local[names[0]].reshape(-1)
            if distributed:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat = flat.cpu()
            off = 0
            for n in names:
                k = local[n].numel()
                out[n] = flat[off:off + k].reshape(local[n].shape).clone()
                off += k
        return out

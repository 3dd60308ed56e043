# -*- coding: utf-8 -*-
"""`deeplab_merge_batch` with the reference's signature (utils/panoptic_merge.py:18-40),
executed by `npb_deeplab_merge` (csrc/merge.cu) for arbitrary semantic / instance /
foreground maps.  (Inside `PanopticPostprocessing` the merge is fused with the grouping
pass instead, see model/postprocessing/panoptic.py.)"""
from ctypes import c_int, c_int64
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib


def deeplab_merge_batch(
    semantic_batch: torch.Tensor,
    instance_batch: torch.Tensor,
    instance_fg_batch: torch.Tensor,
    max_instances_per_category: int,
    thing_ids: Sequence[int],
    void_label: int
) -> Tuple[torch.Tensor, List[Dict[int, int]]]:
    """Returns (panoptic ids (B,H,W) int64 on the inputs' CUDA device, list of
    {panoptic id: instance id}).  `semantic_batch`: any integer dtype, 0 = void;
    `instance_batch`: uint8 ids (0 = none); `instance_fg_batch`: bool."""
    if not semantic_batch.is_cuda:
        raise RuntimeError('deeplab_merge_batch: expected CUDA tensors (no CPU path)')
    dev = semantic_batch.device
    sem = _lib.require_cuda(semantic_batch.to(torch.int64), 'semantic_batch')
    if instance_batch.dtype != torch.uint8:
        if instance_batch.numel() and int(instance_batch.max()) > 255:
            raise ValueError('deeplab_merge_batch: instance ids must be < 256')
        instance_batch = instance_batch.to(torch.uint8)
    ins = _lib.require_cuda(instance_batch.to(dev), 'instance_batch')
    fg = _lib.require_cuda(instance_fg_batch.to(dev).to(torch.uint8), 'instance_fg_batch')
    B = sem.shape[0]
    P = sem[0].numel()
    thing_ids = [int(t) for t in thing_ids]
    n_classes = max([int(sem.max()) + 1 if sem.numel() else 1] + [t + 1 for t in thing_ids])
    lut = _lib.host_lut([c in thing_ids for c in range(n_classes)], n_classes)
    L = _lib.lib()
    ws = torch.empty(L.npb_deeplab_merge_workspace_bytes(B, n_classes), dtype=torch.uint8,
                     device=dev)
    pan = torch.empty(sem.shape, dtype=torch.int64, device=dev)
    inst_class = torch.empty((B, _lib.MAX_INST), dtype=torch.int32, device=dev)
    inst_pan = torch.empty((B, _lib.MAX_INST), dtype=torch.int64, device=dev)
    inst_area = torch.empty((B, _lib.MAX_INST), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.npb_deeplab_merge(
        _lib.ptr(sem), _lib.ptr(ins), _lib.ptr(fg), c_int(B), c_int64(P), c_int(n_classes),
        c_int64(max_instances_per_category), lut, c_int64(void_label), _lib.ptr(ws),
        _lib.ptr(pan), _lib.ptr(inst_class), _lib.ptr(inst_pan), _lib.ptr(inst_area),
        _lib.ptr(status), _lib.stream_ptr(dev)), 'deeplab_merge_batch')
    # numpy on the host copies: indexing tensor elements one by one costs microseconds each
    cls_h, pan_h, status_h = inst_class.cpu().numpy(), inst_pan.cpu().numpy(), status.cpu()
    _lib.raise_for_status(status_h.tolist(), 'deeplab_merge_batch')
    ids = []
    for b in range(B):
        used = np.flatnonzero(cls_h[b, 1:] >= 0) + 1          # ascending instance id
        ids.append(dict(zip(pan_h[b, used].tolist(), used.tolist())))
    return pan, ids


PART_CAP = 4096     # (instance, class) parts per frame (kMaxParts in csrc/merge.cu)


def naive_merge_semantic_and_instance_batch(
    semantic_batch: torch.Tensor,
    instance_batch: torch.Tensor,
    max_instances_per_category: int,
    thing_ids: Sequence[int],
    void_label: int
) -> Tuple[torch.Tensor, List[Dict[int, int]]]:
    """Ground-truth panoptic targets for a batch: the batched GPU form of
    `naive_merge_semantic_and_instance_np` (utils/panoptic_merge.py:43-107), which
    `PanopticTargetGenerator` (data/preprocessing/panoptic.py:16-85) applies per sample.
    `semantic_batch` (B,H,W) integer classes in [0, 255] (0 = void), `instance_batch` (B,H,W)
    integer ids in [0, 65535].  Returns (panoptic ids int64 (B,H,W), list of
    {panoptic id: instance id}) -- `panoptic` and `panoptic_ids_to_instance_dict` of a batch."""
    if not semantic_batch.is_cuda:
        raise RuntimeError('naive_merge_semantic_and_instance_batch: expected CUDA tensors')
    dev = semantic_batch.device
    sem = _lib.require_cuda(semantic_batch.to(torch.uint8), 'semantic_batch', ndim=3)
    ins = _lib.require_cuda(instance_batch.to(dev).to(torch.int32), 'instance_batch', ndim=3)
    B = sem.shape[0]
    P = sem[0].numel()
    thing_ids = [int(t) for t in thing_ids]
    n_classes = 256
    lut = _lib.host_lut([c in thing_ids for c in range(n_classes)], n_classes)
    L = _lib.lib()
    ws = torch.empty(L.npb_naive_merge_workspace_bytes(B), dtype=torch.uint8, device=dev)
    pan = torch.empty(sem.shape, dtype=torch.int64, device=dev)
    part_keys = torch.empty((B, PART_CAP), dtype=torch.int32, device=dev)
    part_pan = torch.empty((B, PART_CAP), dtype=torch.int64, device=dev)
    n_parts = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.npb_naive_merge(
        _lib.ptr(sem), _lib.ptr(ins), c_int(B), c_int64(P), c_int64(max_instances_per_category),
        lut, c_int(n_classes), c_int64(void_label), _lib.ptr(ws), _lib.ptr(pan),
        _lib.ptr(part_keys), _lib.ptr(part_pan), _lib.ptr(n_parts), _lib.ptr(status),
        _lib.stream_ptr(dev)), 'npb_naive_merge')
    _lib.raise_for_status(status.cpu().tolist(), 'naive_merge_semantic_and_instance_batch')
    n_h = n_parts.cpu().tolist()
    n_max = max(n_h, default=0)
    keys_h, pan_h = part_keys[:, :n_max].cpu().tolist(), part_pan[:, :n_max].cpu().tolist()
    ids = [{pan_h[b][t]: (keys_h[b][t] >> 16) & 0xffff for t in range(n_h[b])}
           for b in range(B)]
    return pan, ids

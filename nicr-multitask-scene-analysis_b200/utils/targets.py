# -*- coding: utf-8 -*-
"""Ground-truth instance targets on the GPU, batched.

`InstanceTargetGenerator` mirrors the reference's pre-processing step of the same name
(data/preprocessing/instance.py:97-286), which encodes ONE sample with numpy inside the data
loader: centre heat-map (Gaussian of `sigma` stamped at the centre of mass of every thing
instance), offsets to that centre, foreground and centre masks.  Here a whole batch of
ground-truth maps that already lives on the device is encoded by `npb_instance_targets`
(csrc/targets.cu).
"""
from ctypes import c_int
from typing import Dict, List, Sequence

import numpy as np
import torch

from .. import _lib

LIST_CAP = 4096


class InstanceTargetGenerator:
    def __init__(self, sigma: int, semantic_classes_is_thing: Sequence[bool],
                 normalized_offset: bool = True) -> None:
        """`semantic_classes_is_thing`: one flag per semantic label INCLUDING void (index 0)."""
        self._sigma = int(sigma)
        self._is_thing = tuple(bool(t) for t in semantic_classes_is_thing)
        assert 1 <= len(self._is_thing) <= 256
        self._normalized_offset = bool(normalized_offset)
        self._gauss_host = self._precompute_2d_gauss(self._sigma)
        self._gauss = {}

    @staticmethod
    def _precompute_2d_gauss(sigma: int) -> np.ndarray:
        """(6*sigma+3)^2 stamp, evaluated in float64 exactly like instance.py:140-147 and
        rounded to float32 the way the assignment into the float32 heat-map does."""
        size = 6 * sigma + 3
        x = np.arange(0, size, 1, float)
        y = x[:, np.newaxis]
        c = 3 * sigma + 1
        return np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2)).astype(np.float32)

    def __call__(self, semantic: torch.Tensor, instance: torch.Tensor) -> Dict[str, object]:
        """semantic (B,H,W) integer labels with void = 0, instance (B,H,W) integer ids in
        [0, 65535] (stuff pixels must carry id 0).  Returns `instance_center` (B,H,W) f32,
        `instance_offset` (B,2,H,W) f32 (normalised) or int16 (pixels), `instance_foreground`,
        `instance_center_mask` (B,H,W) bool, and per frame the sorted ids of the
        `encoded_instances` / `skipped_instances_due_to_stuff`."""
        if not semantic.is_cuda:
            raise RuntimeError('InstanceTargetGenerator: expected CUDA tensors (no CPU path)')
        dev = semantic.device
        sem = _lib.require_cuda(semantic.to(torch.uint8), 'semantic', ndim=3)
        ins = _lib.require_cuda(instance.to(dev).to(torch.int32), 'instance', ndim=3)
        B, H, W = sem.shape
        if dev not in self._gauss:
            self._gauss[dev] = torch.from_numpy(self._gauss_host).to(dev).contiguous()
        L = _lib.lib()
        ws = torch.empty(L.npb_instance_targets_workspace_bytes(B), dtype=torch.uint8, device=dev)
        center = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        offset = torch.empty((B, 2, H, W), device=dev,
                             dtype=torch.float32 if self._normalized_offset else torch.int16)
        fg = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        cmask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        enc = torch.empty((B, LIST_CAP), dtype=torch.int32, device=dev)
        skip = torch.empty((B, LIST_CAP), dtype=torch.int32, device=dev)
        n_enc = torch.empty(B, dtype=torch.int32, device=dev)
        n_skip = torch.empty(B, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        n_classes = len(self._is_thing)
        _lib.check(L.npb_instance_targets(
            _lib.ptr(sem), _lib.ptr(ins), c_int(B), c_int(H), c_int(W),
            _lib.host_lut(self._is_thing, n_classes), c_int(n_classes), c_int(self._sigma),
            _lib.ptr(self._gauss[dev]), c_int(int(self._normalized_offset)), _lib.ptr(ws),
            _lib.ptr(center), _lib.ptr(offset), _lib.ptr(fg), _lib.ptr(cmask), _lib.ptr(enc),
            _lib.ptr(skip), _lib.ptr(n_enc), _lib.ptr(n_skip), c_int(LIST_CAP), _lib.ptr(status),
            _lib.stream_ptr(dev)), 'npb_instance_targets')
        code = int(status.item())
        if code == _lib.ERR_ARG:
            # the reference's `assert (instance_image[~foreground] == 0).all()` (instance.py:260)
            raise AssertionError('stuff pixels carry instance ids: clear them first '
                                 '(InstanceClearStuffIDs)')
        _lib.raise_for_status([code], 'InstanceTargetGenerator')
        ne, ns, enc_h, skip_h = n_enc.cpu().tolist(), n_skip.cpu().tolist(), enc.cpu(), skip.cpu()
        encoded: List[List[int]] = [sorted(enc_h[b, :ne[b]].tolist()) for b in range(B)]
        skipped: List[List[int]] = [sorted(skip_h[b, :ns[b]].tolist()) for b in range(B)]
        return {'instance_center': center, 'instance_offset': offset,
                'instance_foreground': fg.view(torch.bool),
                'instance_center_mask': cmask.view(torch.bool),
                'encoded_instances': encoded, 'skipped_instances_due_to_stuff': skipped}

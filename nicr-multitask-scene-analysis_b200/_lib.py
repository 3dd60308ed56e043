# -*- coding: utf-8 -*-
"""ctypes binding of libnicr_panoptic_b200.so (include/nicr_panoptic_b200.h).

There is deliberately NO fallback: if the CUDA library is missing, or a tensor does not
live on a CUDA device, the call raises.  Torch is used for device memory and streams only.
"""
import ctypes
import os
import threading
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int64, c_size_t,
                    c_void_p)
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NPB_LIB_PATH: an alternative build of the same ABI (A/B measurements); the default is the
# in-tree library next to the sources
LIB_PATH = os.environ.get('NPB_LIB_PATH') or os.path.join(_HERE, 'csrc', 'libnicr_panoptic_b200.so')

MAX_INST = 256
MAX_WIDE_CENTERS = 8192     # NPB_MAX_WIDE_CENTERS

OK = 0
ERR_ARG, ERR_TOO_MANY_CENTERS, ERR_ZERO_DIVISION, ERR_CATEGORY_RANGE, ERR_CAPACITY, ERR_CUDA = \
    -1, -2, -3, -4, -5, -6

U8, I16, I32, I64, BOOL = 0, 1, 2, 3, 4
_DTYPE_CODES = {torch.uint8: U8, torch.int16: I16, torch.int32: I32, torch.int64: I64,
                torch.bool: BOOL}

_P = c_void_p


class EvalArgs(ctypes.Structure):
    """`npb_eval_args` of include/nicr_panoptic_b200.h (evaluation half of the fused calls)."""
    _fields_ = [('target', _P), ('sem_target', _P), ('num_categories', c_int),
                ('confmat_n', c_int), ('ignored_label', c_int64), ('offset', c_int64),
                ('void_segment_id', c_int64), ('workspace', _P), ('iou', _P), ('tp', _P),
                ('fn', _P), ('fp', _P), ('confmat', _P), ('frame_stats', _P), ('matches', _P),
                ('match_cap', c_int), ('n_matches', _P), ('status', _P)]


_FORWARD_ARGS = [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_float,
                 c_int, c_int, c_int, c_int, c_int, c_float, c_int64, _P, _P,
                 _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]
_SIGNATURES = {
    'npb_abi_version': (c_int, []),
    'npb_error_string': (c_char_p, [c_int]),
    'npb_last_cuda_error': (c_char_p, []),
    'npb_build_info': (c_char_p, []),
    'npb_semantic_argmax': (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    'npb_softmax': (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    'npb_thing_mask': (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    'npb_widen_u8': (c_int, [_P, c_int64, c_int64, _P, _P]),
    'npb_resize_nearest': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, _P, _P]),
    'npb_resize_bilinear': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, _P, _P]),
    'npb_semantic_argmax_resized': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_int, _P, _P, _P]),
    'npb_instance_centers_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'npb_instance_centers': (c_int, [_P, c_int, c_int, c_int, c_float, c_int, c_int, _P, c_int,
                                     _P, _P, _P, _P, _P, _P]),
    'npb_group_pixels': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P,
                                 c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    'npb_filter_stuff_area': (c_int, [_P, _P, c_int, c_int64, c_int, c_int64, c_int64, c_int64, _P, _P]),
    'npb_overflow_centers': (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P]),
    'npb_group_pixels_wide': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int,
                                      c_float, _P, _P, _P, _P, _P, _P]),
    'npb_finalize_instances': (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int64, c_int64, _P, _P,
                                       _P, _P, _P, _P]),
    'npb_write_panoptic': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int64, _P,
                                   _P, _P]),
    'npb_panoptic_forward_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'npb_panoptic_forward_workspace_init': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P]),
    'npb_panoptic_forward': (c_int, _FORWARD_ARGS + [_P]),
    'npb_panoptic_forward_eval': (c_int, _FORWARD_ARGS + [POINTER(EvalArgs), _P]),
    'npb_panoptic_forward_eval_pipelined': (c_int, _FORWARD_ARGS + [POINTER(EvalArgs), POINTER(EvalArgs),
                                                            c_int, _P]),
    'npb_pq_match_pending': (c_int, [POINTER(EvalArgs), c_int, c_int64, _P]),
    'npb_write_panoptic_eval': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int64,
                                        _P, _P, POINTER(EvalArgs), _P]),
    'npb_panoptic_scores': (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P,
                                    _P, _P, _P, _P]),
    'npb_deeplab_merge_workspace_bytes': (c_size_t, [c_int, c_int]),
    'npb_deeplab_merge': (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int64, _P, c_int64, _P, _P,
                                  _P, _P, _P, _P, _P]),
    'npb_naive_merge_workspace_bytes': (c_size_t, [c_int]),
    'npb_naive_merge': (c_int, [_P, _P, c_int, c_int64, c_int64, _P, c_int, c_int64, _P, _P, _P, _P,
                                _P, _P, _P]),
    'npb_instance_targets_workspace_bytes': (c_size_t, [c_int]),
    'npb_instance_targets': (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, _P, _P,
                                     _P, _P, _P, _P, _P, _P, c_int, _P, _P]),
    'npb_instance_orientation': (c_int, [_P, _P, c_int, _P, c_int, c_int64, c_int, _P, _P, _P,
                                         _P, _P]),
    'npb_confmat_update': (c_int, [_P, c_int, _P, c_int, c_int64, c_int, _P, _P, _P]),
    'npb_confmat_update_nonvoid': (c_int, [_P, c_int, _P, c_int, c_int64, c_int, _P, _P, _P]),
    'npb_pq_update_workspace_bytes': (c_size_t, [c_int, c_int]),
    'npb_pq_update_big_frame_workspace_bytes': (c_size_t, [c_int64, c_int]),
    'npb_pq_update_big_frame': (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, _P,
                                        _P, _P, _P, _P, _P, c_int, _P, _P, _P]),
    'npb_pq_update': (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int64, c_int64, c_int64,
                              c_int64, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_switched = threading.local()       # devices to switch back to (see stream_ptr / check)


class NpbError(RuntimeError):
    """Error reported by libnicr_panoptic_b200 (host return code or device status word)."""

    def __init__(self, code: int, where: str = ''):
        self.code = int(code)
        msg = lib().npb_error_string(self.code).decode()
        if self.code == ERR_CUDA:
            msg += ': ' + lib().npb_last_cuda_error().decode()
        super().__init__(f'{where}: {msg} (code {self.code})' if where else
                         f'{msg} (code {self.code})')


def lib():
    """Load the shared library once.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python __graft_entry__.py build` '
                '(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)     # AttributeError if the ABI lost a symbol
            except AttributeError:
                if os.environ.get('NPB_LIB_PATH'):      # an older build in an A/B measurement
                    continue
                raise
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = _Library(handle)
    return _lib


class _Library:
    """The loaded library.  Every entry point that takes a stream is wrapped so that a device
    switch made by `stream_ptr` while the arguments were built is undone when the call returns
    OR raises (e.g. ctypes.ArgumentError during argument conversion)."""

    def __init__(self, handle):
        self._handle = handle
        for name, (_, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                continue
            setattr(self, name, self._guarded(fn) if argtypes and argtypes[-1] is _P and
                    name.startswith('npb_') and 'workspace_bytes' not in name else fn)

    @staticmethod
    def _guarded(fn):
        def call(*args):
            try:
                return fn(*args)
            finally:
                restore_device()
        call.__name__ = fn.__name__
        return call

    def __getattr__(self, name):        # anything not in _SIGNATURES (debug builds: timeline)
        return getattr(self._handle, name)


def raise_for_code(code: int, where: str = '') -> None:
    """Return code of a library call -> exception."""
    if code != OK:
        raise NpbError(code, where)


def check(code: int, where: str = '') -> None:
    """Return code of a library call -> exception.  Also undoes the device switch `stream_ptr`
    made for that call (see there)."""
    restore_device()
    raise_for_code(code, where)


def restore_device() -> None:
    """Undo the device switches of `stream_ptr` of the current thread (idempotent).  `check`
    calls it; call sites that may fail between `stream_ptr` and `check` (argument conversion)
    can call it from a `finally`."""
    stack = getattr(_switched, 'stack', None)
    while stack:
        torch.cuda.set_device(stack.pop())


class on_device:
    """Context manager: the library launches on the CURRENT device of the calling thread, so the
    device of the tensors is made current for the call and restored afterwards -- also when the
    call raises (ctypes.ArgumentError and the like).  With one process per GPU nothing is switched."""
    __slots__ = ('index', 'previous')

    def __init__(self, device):
        device = torch.device(device)
        self.index = device.index
        self.previous = None

    def __enter__(self):
        if self.index is not None:
            current = torch.cuda.current_device()
            if current != self.index:
                self.previous = current
                torch.cuda.set_device(self.index)
        return self

    def __exit__(self, *exc):
        if self.previous is not None:
            torch.cuda.set_device(self.previous)
        return False


def raise_for_status(status, where: str = '') -> None:
    """`status`: iterable of int status words already on the host."""
    worst = min((int(s) for s in status), default=0)
    if worst == ERR_ZERO_DIVISION:
        # the reference's `intersection_area / union` raises exactly this (metric/pq.py:145)
        raise ZeroDivisionError('division by zero')
    if worst != OK:
        raise NpbError(worst, where)


def require_cuda(t: torch.Tensor, name: str, dtype: Optional[torch.dtype] = None,
                 ndim: Optional[int] = None) -> torch.Tensor:
    """Validate a tensor handed to a kernel: CUDA, (dtype), (ndim); returns it contiguous."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name}: expected a torch.Tensor, got {type(t).__name__}')
    if not t.is_cuda:
        raise RuntimeError(f'{name}: expected a CUDA tensor (this implementation has no CPU '
                           f'path); got device {t.device}')
    if dtype is not None and t.dtype != dtype:
        if dtype == torch.float32 and t.dtype in (torch.float16, torch.bfloat16, torch.float64):
            # decoder outputs of a network that ran under autocast (or in double precision): the
            # kernels compute in float32.  Widening half / bfloat16 is exact, so the results are
            # those of the float32 path on the very same values.
            t = t.float()
        else:
            raise TypeError(f'{name}: expected dtype {dtype}, got {t.dtype}')
    if ndim is not None and t.ndim != ndim:
        raise ValueError(f'{name}: expected {ndim} dims, got shape {tuple(t.shape)}')
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device) -> c_void_p:
    """Current stream of `device` as the `stream` argument of a library call.

    The kernels launch on the CURRENT device of the calling thread, so for tensors that live on
    another device of the process (one process driving several GPUs) that device is made current
    here; the call wrapper of `lib()` switches back when the call returns or raises (`check`
    does so as well).  With one process per GPU (the usual set-up) nothing is switched."""
    device = torch.device(device)
    current = torch.cuda.current_device()
    index = current if device.index is None else device.index
    if index != current:
        stack = getattr(_switched, 'stack', None)
        if stack is None:
            stack = _switched.stack = []
        stack.append(current)
        torch.cuda.set_device(index)
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODES[t.dtype]
    except KeyError:
        raise TypeError(f'unsupported integer dtype {t.dtype}') from None


def host_lut(flags: Sequence[bool], n: int):
    """bool sequence -> ctypes uint8 array (host look-up table argument `h_*_lut`); the arrays
    are read-only for the library and cached per (flags, n)."""
    key = (tuple(bool(f) for f in flags), int(n))
    arr = _LUT_CACHE.get(key)
    if arr is None:
        arr = (ctypes.c_uint8 * max(n, 1))()
        for i, f in enumerate(key[0]):
            if i < n and f:
                arr[i] = 1
        if len(_LUT_CACHE) > 256:
            _LUT_CACHE.clear()
        _LUT_CACHE[key] = arr
    return arr


_LUT_CACHE = {}

# -*- coding: utf-8 -*-
"""Containers for post-processing results.

`ResultDict` is a plain `dict` whose values may be *deferred*: a few entries of the
reference's result dict are expensive, rarely read copies of information the kernels keep
in compact form (e.g. the (B,C,H,W) soft-max scores, int64 twins of uint8 maps).  They are
present as keys (so `k in result`, `result.keys()` behave like the reference's dict) and are
materialised by the owning kernel call on first access.

`InstanceTables` holds the per-instance tables of a batch after their single device->host
copy and turns them into the python dicts / lists of the reference API.
"""
from typing import Any, Callable, Dict, List

import numpy as np
import torch

from . import _lib


class _Deferred:
    """`fn()` -- or `fn(result_dict)` when `wants_dict` -- produces the value on first access.
    Entries that need the dict itself get it handed in at that moment instead of closing over it:
    a closure over the dict would make it a reference cycle, and its tensors (tens of megabytes
    per batch) would stay allocated until the garbage collector gets round to it."""
    __slots__ = ('fn', 'wants_dict')

    def __init__(self, fn: Callable[..., Any], wants_dict: bool = False):
        self.fn = fn
        self.wants_dict = wants_dict


class _Alias:
    __slots__ = ('source',)

    def __init__(self, source: str):
        self.source = source


class ResultDict(dict):
    def defer(self, key: str, fn: Callable[[], Any]) -> None:
        dict.__setitem__(self, key, _Deferred(fn))

    def defer_with_dict(self, key: str, fn: Callable[['ResultDict'], Any]) -> None:
        """like `defer`, `fn` receives this dict"""
        dict.__setitem__(self, key, _Deferred(fn, wants_dict=True))

    def alias(self, key: str, source: str) -> None:
        """`key` resolves to whatever `source` resolves to (identity full-res twins)."""
        dict.__setitem__(self, key, _Alias(source))

    def _resolve(self, key, value):
        if isinstance(value, _Deferred):
            value = value.fn(self) if value.wants_dict else value.fn()
            dict.__setitem__(self, key, value)
        elif isinstance(value, _Alias):
            value = self[value.source]
            dict.__setitem__(self, key, value)
        return value

    def __getitem__(self, key):
        return self._resolve(key, dict.__getitem__(self, key))

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        if key in self:
            value = self[key]
            dict.__delitem__(self, key)
            return value
        if default:
            return default[0]
        raise KeyError(key)

    def items(self):
        return [(k, self[k]) for k in list(dict.keys(self))]

    def values(self):
        return [self[k] for k in list(dict.keys(self))]

    def is_deferred(self, key: str) -> bool:
        return isinstance(dict.__getitem__(self, key), (_Deferred, _Alias))

    def materialize(self) -> 'ResultDict':
        for k in list(dict.keys(self)):
            self[k]
        return self


class InstanceTables:
    """Per-instance tables of one batch (rows of `_lib.MAX_INST`, row 0 unused).

    One packed device buffer -> ONE device->host copy (a few hundred KB), done on first
    access by `wait()`, which also checks the per-frame status words and exposes numpy views.
    """

    FIELDS = (('status', np.int32, 1), ('n_centers', np.int32, 1),
              ('centers_yx', np.int32, 2 * _lib.MAX_INST), ('center_score', np.float32, _lib.MAX_INST),
              ('inst_class', np.int32, _lib.MAX_INST), ('inst_area', np.int32, _lib.MAX_INST),
              ('inst_angle', np.float32, _lib.MAX_INST), ('inst_pan_id', np.int64, _lib.MAX_INST))

    _layouts: Dict[int, Any] = {}

    @classmethod
    def layout(cls, batch_size: int):
        """byte offsets of the fields inside the packed buffer (largest alignment first)"""
        cached = cls._layouts.get(batch_size)
        if cached is None:
            cached = cls._layouts[batch_size] = cls._layout(batch_size)
        return cached

    @classmethod
    def _layout(cls, batch_size: int):
        offsets, off = {}, 0
        for name, dt, per_frame in sorted(cls.FIELDS, key=lambda f: -np.dtype(f[1]).itemsize):
            nbytes = np.dtype(dt).itemsize * per_frame * batch_size
            offsets[name] = (off, nbytes, dt, per_frame)
            off += (nbytes + 15) // 16 * 16
        return offsets, off

    def __init__(self, batch_size: int, device: torch.device, storage: torch.Tensor = None):
        """`storage`: optional uint8 device tensor of `layout(B)[1]` bytes to live in (the
        status words inside it must be zero on entry of the kernels that report into it)."""
        self.B = batch_size
        self.device = device
        self._offsets, self.nbytes = self.layout(batch_size)
        self.dev = storage if storage is not None else \
            torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self._np = None
        self._host = None           # (pinned copy, event) once prefetch() ran
        self._where = 'panoptic post-processing'
        # frames redone with more than 255 centres (`on_overflow='wrap'`, instance.py:236):
        # frame -> (centres (K,2) int32, scores (K,) float32) on the host; their table rows hold
        # the 255 wrapped instance ids
        self.wide: Dict[int, Any] = {}

    def dptr(self, name: str):
        off, _, _, _ = self._offsets[name]
        import ctypes
        return ctypes.c_void_p(self.dev.data_ptr() + off)

    def dptr_row(self, name: str, b: int):
        """device pointer of frame `b`'s row of a table"""
        off, _, dt, per_frame = self._offsets[name]
        import ctypes
        return ctypes.c_void_p(self.dev.data_ptr() + off + b * per_frame * np.dtype(dt).itemsize)

    def dview(self, name: str) -> torch.Tensor:
        off, nbytes, dt, per_frame = self._offsets[name]
        tdt = {np.int32: torch.int32, np.float32: torch.float32, np.int64: torch.int64}[dt]
        v = self.dev[off:off + nbytes].view(tdt)
        return v.view(self.B, per_frame) if per_frame > 1 else v

    def prefetch(self) -> 'InstanceTables':
        """Start the device->host copy of the tables on the CURRENT stream into pinned memory;
        `wait()` then only waits for this copy, not for whatever was enqueued afterwards."""
        pinned = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(self.dev, non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(self.device))
        self._host = (pinned, event)
        self._np = None
        return self

    # pinned staging buffers for the synchronous download (cudaHostAlloc costs tens of
    # microseconds: they are kept); the tables are copied out of them before they are reused
    _staging: Dict[Any, List[torch.Tensor]] = {}

    def wait(self) -> 'InstanceTables':
        if self._np is None:
            if self._host is not None:
                self._host[1].synchronize()
                raw = self._host[0].numpy()
            else:
                pool = self._staging.setdefault((self.nbytes, self.device), [])
                pinned = pool.pop() if pool else torch.empty(self.nbytes, dtype=torch.uint8,
                                                             pin_memory=True)
                pinned.copy_(self.dev, non_blocking=True)
                # blocks until the producing kernels and this copy are done
                torch.cuda.current_stream(self.device).synchronize()
                raw = pinned.numpy().copy()
                if len(pool) < 4:
                    pool.append(pinned)
            self._np = {}
            for name, (off, nbytes, dt, per_frame) in self._offsets.items():
                a = raw[off:off + nbytes].view(dt)
                self._np[name] = a.reshape(self.B, per_frame) if per_frame > 1 else a
            _lib.raise_for_status(self._np['status'], self._where)
        return self

    def __getitem__(self, name: str) -> np.ndarray:
        return self.wait()._np[name]

    def invalidate(self) -> None:
        """Forget the host copy: the device buffer was rewritten (CUDA-graph replay)."""
        self._np = None
        self._host = None

    # ---- python structures of the reference API ------------------------------------------
    def _rows(self):
        """(n per frame as a list, largest n): the python structures only look at the used rows
        of the 256-row tables"""
        n = self['n_centers'].tolist()
        return n, (max(n) if n else 0)

    def centers_list(self) -> List[torch.Tensor]:
        """instance.py:163-166: list of (n, 2) int32 tensors (y, x), raster order."""
        n = self['n_centers']
        c = self['centers_yx'].reshape(self.B, _lib.MAX_INST, 2)
        return [torch.from_numpy(self.wide[b][0].copy() if b in self.wide else c[b, :n[b]].copy())
                for b in range(self.B)]

    def meta(self, with_orientation: bool = False) -> List[Dict[int, Dict[str, Any]]]:
        """instance.py:253-266: {id: {'center_yx', 'area', 'score'}} for every centre;
        `with_orientation`: + 'orientation' (NaN for instances without one, panoptic.py:305-314)."""
        n, m = self._rows()
        # columns, not (y, x) pairs: one list per frame and coordinate instead of one per centre
        # (every container allocated here counts towards the thresholds of python's cyclic GC)
        c = self['centers_yx'].reshape(self.B, _lib.MAX_INST, 2)
        ys, xs = c[:, :m, 0].tolist(), c[:, :m, 1].tolist()
        area = self['inst_area'][:, :m + 1].tolist()
        score = self['center_score'][:, :m].tolist()
        if not with_orientation:
            out = [{i + 1: {'center_yx': (yb[i], xb[i]), 'area': ab[i + 1], 'score': sb[i]}
                    for i in range(nb)} for nb, yb, xb, ab, sb in zip(n, ys, xs, area, score)]
        else:
            ang = self['inst_angle'][:, :m + 1].tolist()
            out = [{i + 1: {'center_yx': (yb[i], xb[i]), 'area': ab[i + 1], 'score': sb[i],
                            'orientation': gb[i + 1]}
                    for i in range(nb)} for nb, yb, xb, ab, sb, gb in zip(n, ys, xs, area, score, ang)]
        for b in self.wide:
            out[b] = self._meta_wide(b, with_orientation)
        return out

    def _meta_wide(self, b: int, with_orientation: bool) -> Dict[int, Dict[str, Any]]:
        """Meta dict of a frame with K > 255 centres the way the reference builds it
        (instance.py:253-266): one entry per centre; areas are the bincount of the WRAPPED uint8
        ids, so entries beyond 255 have area 0 (and no orientation: NaN, panoptic.py:311-314)."""
        cyx, score = self.wide[b]
        area = self['inst_area'][b].tolist()
        ang = self['inst_angle'][b].tolist() if with_orientation else None
        nan = float('nan')
        out = {}
        for i, ((y, x), sc) in enumerate(zip(cyx.tolist(), score.tolist()), start=1):
            e = {'center_yx': (y, x), 'area': area[i] if i < _lib.MAX_INST else 0, 'score': sc}
            if with_orientation:
                e['orientation'] = ang[i] if i < _lib.MAX_INST else nan
            out[i] = e
        return out

    def panoptic_ids(self) -> List[Dict[int, int]]:
        """panoptic_merge.py:209: {panoptic id: raw instance id}, ascending instance id."""
        n, m = self._rows()
        cls = self['inst_class'][:, :m + 1].tolist()
        pan = self['inst_pan_id'][:, :m + 1].tolist()
        return [{pb[i]: i for i in range(1, nb + 1) if cb[i] >= 0}
                for nb, cb, pb in zip(n, cls, pan)]

    def orientations(self) -> List[Dict[int, float]]:
        """instance.py:301-317: {raw instance id: angle} for instances whose panoptic class
        carries an orientation (angle is NaN-free there by construction)."""
        n, m = self._rows()
        ang = self['inst_angle'][:, :m + 1].tolist()
        return [{i: ab[i] for i in range(1, nb + 1) if ab[i] == ab[i]}      # NaN != NaN
                for nb, ab in zip(n, ang)]

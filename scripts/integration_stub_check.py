#!/usr/bin/env python
"""Runs the reference-side ctypes stubs of INTEGRATION.md section 2 (metric side) as written
there, against this package's own metric classes on a small random batch."""
import ctypes
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = ctypes.CDLL(os.path.join(ROOT, 'nicr-multitask-scene-analysis_b200', 'csrc',
                                'libnicr_panoptic_b200.so'))
_p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None

# ---- verbatim from INTEGRATION.md ---------------------------------------------------------
_U8, _I16, _I32, _I64, _BOOL = range(5)            # NPB_U8 ... NPB_BOOL
_CODE = {torch.uint8: _U8, torch.int16: _I16, torch.int32: _I32, torch.int64: _I64, torch.bool: _BOOL}
_stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def miou_update(self, preds, target):               # metric/miou.py:44-56
    preds, target = preds.contiguous(), target.contiguous()
    status = torch.zeros(1, dtype=torch.int32, device=preds.device)
    rc = _lib.npb_confmat_update(_p(preds), _CODE[preds.dtype], _p(target), _CODE[target.dtype],
                                 ctypes.c_int64(preds.numel()), self._n_classes,
                                 _p(self.confmat), _p(status), _stream())
    assert rc == 0 and int(status) == 0             # values outside [0, n_classes)

def pq_update(self, preds, targets):                # metric/pq.py:264-303, no process pool
    B, P = preds.shape[0], preds[0].numel()
    _lib.npb_pq_update_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(_lib.npb_pq_update_workspace_bytes(B, self.num_categories),
                     dtype=torch.uint8, device=preds.device)
    status = torch.zeros(B, dtype=torch.int32, device=preds.device)
    i64 = ctypes.c_int64
    rc = _lib.npb_pq_update(
        _p(preds.contiguous()), _p(targets.contiguous()), None, B, i64(P), self.num_categories,
        i64(self.ignored_label), i64(self.max_instances_per_category), i64(self.offset),
        i64(self.void_segment_id), _p(ws), _p(self.iou_per_class), _p(self.tp_per_class),
        _p(self.fn_per_class), _p(self.fp_per_class), None, 0, None, None, 0, None, _p(status),
        _stream())
    assert rc == 0
    assert int((status != 0).sum()) == 0
# -------------------------------------------------------------------------------------------


def main():
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticQuality
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(0)
    NC, L, OFF, B, H, W = 6, 1 << 16, 256 ** 3, 3, 40, 56
    low = torch.randint(0, NC, (B, 5, 7), generator=g)
    cat = low.repeat_interleave(8, 1).repeat_interleave(8, 2)
    inst = torch.randint(1, 3, (B, 5, 7), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    is_thing = [False, True, False, True, True, False]
    tgt = (cat * L + torch.where(torch.tensor(is_thing)[cat], inst, torch.zeros_like(inst))).to(dev)
    pred = torch.roll(tgt, 3, -1).contiguous()
    ours_pq = PanopticQuality(NC, 0, L, OFF, is_thing, device=dev)
    ours_pq.update(pred, tgt)
    ours_pq.check_status()
    stub = SimpleNamespace(num_categories=NC, ignored_label=0, max_instances_per_category=L, offset=OFF,
                           void_segment_id=0,
                           **{k: torch.zeros(NC, dtype=torch.float64, device=dev)
                              for k in ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class')})
    pq_update(stub, pred, tgt)
    for k in ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class'):
        assert torch.equal(getattr(stub, k), getattr(ours_pq, k)), k
    ours_m = MeanIntersectionOverUnion(NC, device=dev)
    ours_m.update(pred // L, (tgt // L).to(torch.uint8))
    stub_m = SimpleNamespace(_n_classes=NC, confmat=torch.zeros((NC, NC), dtype=torch.int64, device=dev))
    miou_update(stub_m, pred // L, (tgt // L).to(torch.uint8))
    assert torch.equal(stub_m.confmat, ours_m.confmat)
    print('integration stubs ok: tp', stub.tp_per_class.tolist(), 'pixels', int(stub_m.confmat.sum()))


if __name__ == '__main__':
    main()

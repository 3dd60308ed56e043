"""CPU ORACLE -- test infrastructure, NOT product code.

numpy/ctypes front-end of `panoptic_oracle.c`, the plain-C restatement of the
reference's panoptic post-processing + evaluation path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this package;
`nicr_mt_scene_analysis_b200` never does (tests/test_no_oracle_in_product.py checks).

Parity status: PINNED against the live reference (tests/golden/, see make_golden.py).
"""
import ctypes
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_int64, c_long, c_uint8, c_void_p
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import build as _build

_LIB = None

ERRORS = {-1: 'bad argument', -2: 'more than 255 centres in a frame',
          -3: 'ZeroDivisionError (union == 0)', -4: 'category/class id out of range',
          -5: 'output capacity exceeded'}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f'oracle error {code}: {ERRORS.get(code, "?")}')
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_build.build())
    return _LIB


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _check(code):
    if code != 0:
        raise OracleError(code)


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(c_int(n))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def semantic_argmax(logits: np.ndarray) -> np.ndarray:
    """(B,C,H,W) f32 -> (B,H,W) u8, first-index arg-max (semantic.py:52-53)."""
    logits = _c(logits, np.float32)
    B, C, H, W = logits.shape
    out = np.empty((B, H, W), np.uint8)
    _check(lib().orc_semantic_argmax(_p(logits), c_int(B), c_int(C), c_long(H * W), _p(out)))
    return out


def semantic_score(logits: np.ndarray) -> np.ndarray:
    logits = _c(logits, np.float32)
    B, C, H, W = logits.shape
    out = np.empty((B, H, W), np.float32)
    _check(lib().orc_semantic_score(_p(logits), c_int(B), c_int(C), c_long(H * W), _p(out)))
    return out


def instance_centers(heat, threshold=0.1, nms_kernel_size=3, top_k=64, foreground=None,
                     apply_foreground_mask=False, cap=255
                     ) -> Tuple[np.ndarray, List[np.ndarray]]:
    """instance.py:78-168 -> (bool map (B,H,W), list of (n,2) int32 (y,x))."""
    heat = _c(heat, np.float32)
    B, H, W = heat.shape[0], heat.shape[-2], heat.shape[-1]
    fg = None if foreground is None else _c(foreground, np.uint8)
    centers = np.zeros((B, cap, 2), np.int32)
    n = np.zeros((B,), np.int32)
    mask = np.zeros((B, H, W), np.uint8)
    _check(lib().orc_instance_centers(
        _p(heat), c_int(B), c_int(H), c_int(W), c_float(threshold), c_int(nms_kernel_size),
        c_int(top_k), _p(fg), c_int(int(apply_foreground_mask)), _p(centers), c_int(cap),
        _p(n), _p(mask)))
    return mask.astype(bool), [centers[b, :n[b]].copy() for b in range(B)]


class allow_wrap:
    """`with oracle.allow_wrap():` -- more than 255 centres wrap like the reference's uint8 ids
    (instance.py:236) instead of being rejected; pass `cap` >= the number of centres."""

    def __enter__(self):
        lib().orc_set_allow_wrap(c_int(1))

    def __exit__(self, *exc):
        lib().orc_set_allow_wrap(c_int(0))


def instance_segmentation(heat, offset, foreground, threshold=0.1, nms_kernel_size=3,
                          top_k=64, apply_foreground_mask=False, normalized_offset=True,
                          offset_distance_threshold=None, cap=255):
    """instance.py:170-268 (+ panoptic.py:105-111 offset de-normalisation when
    `normalized_offset`).  Returns (inst u8 (B,H,W), meta list of dicts)."""
    heat = _c(heat, np.float32)
    offset = _c(offset, np.float32)
    fg = _c(foreground, np.uint8)
    B, H, W = heat.shape[0], heat.shape[-2], heat.shape[-1]
    inst = np.zeros((B, H, W), np.uint8)
    centers = np.zeros((B, cap, 2), np.int32)
    n = np.zeros((B,), np.int32)
    area = np.zeros((B, cap + 1), np.int32)
    score = np.zeros((B, cap), np.float32)
    use_thr = offset_distance_threshold is not None
    _check(lib().orc_instance_segmentation(
        _p(heat), _p(offset), _p(fg), c_int(B), c_int(H), c_int(W), c_float(threshold),
        c_int(nms_kernel_size), c_int(top_k), c_int(int(apply_foreground_mask)),
        c_int(int(normalized_offset)), c_int(int(use_thr)),
        c_float(offset_distance_threshold if use_thr else 0.0), _p(inst), _p(centers),
        c_int(cap), _p(n), _p(area), _p(score)))
    meta = []
    for b in range(B):
        meta.append({i + 1: {'center_yx': (int(centers[b, i, 0]), int(centers[b, i, 1])),
                             'area': int(area[b, i + 1]),
                             'score': float(score[b, i])} for i in range(n[b])})
    return inst, meta


def deeplab_merge_batch(semantic, instance, instance_fg, max_instances_per_category,
                        thing_ids, void_label) -> Tuple[np.ndarray, List[Dict[int, int]]]:
    """utils/panoptic_merge.py:18-40 / 172-225."""
    sem = _c(semantic, np.int32)
    ins = _c(instance, np.uint8)
    fg = _c(instance_fg, np.uint8)
    B = sem.shape[0]
    P = int(np.prod(sem.shape[1:]))
    n_classes = int(max(int(sem.max()) + 1, max(list(thing_ids) + [0]) + 1))
    thing = np.zeros((n_classes,), np.uint8)
    thing[np.asarray(list(thing_ids), dtype=np.int64)] = 1
    pan = np.empty(sem.shape, np.int64)
    pairs = np.zeros((B, 256, 2), np.int64)
    n_pairs = np.zeros((B,), np.int32)
    _check(lib().orc_deeplab_merge_batch(
        _p(sem), _p(ins), _p(fg), c_int(B), c_long(P), c_int(n_classes),
        c_int64(max_instances_per_category), _p(thing), c_int64(void_label), _p(pan),
        _p(pairs), _p(n_pairs)))
    dicts = [{int(pairs[b, i, 0]): int(pairs[b, i, 1]) for i in range(n_pairs[b])}
             for b in range(B)]
    return pan, dicts


def naive_merge_batch(semantic, instance, max_instances_per_category, thing_ids, void_label,
                      cap=4096) -> Tuple[np.ndarray, List[Dict[int, int]]]:
    """utils/panoptic_merge.py:43-107 for a batch."""
    sem = _c(semantic, np.uint8)
    ins = _c(instance, np.int32)
    B = sem.shape[0]
    P = int(np.prod(sem.shape[1:]))
    thing = np.zeros((256,), np.uint8)
    thing[np.asarray(list(thing_ids), dtype=np.int64)] = 1
    pan = np.empty(sem.shape, np.int64)
    pairs = np.zeros((B, cap, 2), np.int64)
    n_pairs = np.zeros((B,), np.int32)
    _check(lib().orc_naive_merge_batch(_p(sem), _p(ins), c_int(B), c_long(P),
                                       c_int64(max_instances_per_category), _p(thing),
                                       c_int64(void_label), _p(pan), _p(pairs), c_int(cap),
                                       _p(n_pairs)))
    return pan, [{int(pairs[b, i, 0]): int(pairs[b, i, 1]) for i in range(n_pairs[b])}
                 for b in range(B)]


def gauss_stamp(sigma: int) -> np.ndarray:
    """instance.py:140-147, float64 -> float32."""
    size = 6 * sigma + 3
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    c = 3 * sigma + 1
    return np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2)).astype(np.float32)


def instance_targets(semantic, instance, sigma, is_thing_with_void, normalized_offset=True):
    """data/preprocessing/instance.py:151-286 for a batch -> dict of arrays, offset (B,2,H,W)."""
    sem = _c(semantic, np.uint8)
    ins = _c(instance, np.int32)
    B, H, W = sem.shape
    thing = np.zeros((256,), np.uint8)
    thing[:len(is_thing_with_void)] = np.asarray(is_thing_with_void, np.uint8)
    gauss = np.ascontiguousarray(gauss_stamp(sigma))
    center = np.empty((B, H, W), np.float32)
    offset = np.empty((B, 2, H, W), np.float32)
    fg = np.empty((B, H, W), np.uint8)
    cmask = np.empty((B, H, W), np.uint8)
    code = lib().orc_instance_targets(_p(sem), _p(ins), c_int(B), c_int(H), c_int(W), _p(thing),
                                      c_int(sigma), _p(gauss), c_int(int(normalized_offset)),
                                      _p(center), _p(offset), _p(fg), _p(cmask))
    if code == -1:
        raise AssertionError('stuff pixels carry instance ids')
    _check(code)
    return {'instance_center': center, 'instance_offset': offset,
            'instance_foreground': fg.astype(bool), 'instance_center_mask': cmask.astype(bool)}


def instance_orientation(orientation, instance_segmentation, foreground_mask=None
                         ) -> List[Dict[int, float]]:
    """instance.py:270-319."""
    ori = _c(orientation, np.float32)
    seg = _c(instance_segmentation, np.int32)
    B = ori.shape[0]
    P = int(np.prod(seg.shape[1:]))
    mask = None if foreground_mask is None else _c(foreground_mask, np.uint8)
    max_id = max(int(seg.max()), 1)
    present = np.zeros((B, max_id + 1), np.uint8)
    angle = np.zeros((B, max_id + 1), np.float32)
    sums = np.zeros((B, max_id + 1, 2), np.float64)
    _check(lib().orc_instance_orientation(_p(ori), _p(seg), _p(mask), c_int(B), c_long(P),
                                          c_int(max_id), _p(present), _p(angle), _p(sums)))
    return [{i: float(angle[b, i]) for i in range(1, max_id + 1) if present[b, i]}
            for b in range(B)]


def confmat(preds, target, n_classes: int) -> np.ndarray:
    """miou.py:44-56: rows = target, cols = pred, int64."""
    p = _c(preds, np.int64).reshape(-1)
    t = _c(target, np.int64).reshape(-1)
    cm = np.zeros((n_classes, n_classes), np.int64)
    _check(lib().orc_confmat(_p(p), _p(t), c_long(p.size), c_int(n_classes), _p(cm)))
    return cm


def miou_from_confmat(cm: np.ndarray, ignore_first_class: bool):
    """miou.py:58-94 (float32 arithmetic like the reference)."""
    cm = np.asarray(cm, np.int64)
    n = cm.shape[0]
    tp = np.diag(cm).astype(np.float32)
    sum_pred = cm.sum(0).astype(np.float32)
    sum_gt = cm.sum(1).astype(np.float32)
    if ignore_first_class:
        tp, sum_pred, sum_gt = tp[1:], sum_pred[1:], sum_gt[1:]
        sum_pred = sum_pred - cm[0, 1:].astype(np.float32)
    mask = sum_gt != 0
    iou = tp[mask] / (sum_pred[mask] + sum_gt[mask] - tp[mask])
    ious = np.full((n,), np.nan, np.float32)
    idx = np.nonzero(mask)[0] + (1 if ignore_first_class else 0)
    ious[idx] = iou
    return np.float32(iou.mean()) if iou.size else np.float32(np.nan), ious


def pq_compare_and_accumulate(pred, target, num_categories, ignored_label,
                              max_instances_per_category, offset, void_segment_id, cap=4096):
    """pq.py:60-179 for one frame -> (iou, tp, fn, fp) f64 + set of (gt_id, pred_id)."""
    p = _c(pred, np.int64).reshape(-1)
    t = _c(target, np.int64).reshape(-1)
    iou, tp, fn, fp = (np.zeros((num_categories,), np.float64) for _ in range(4))
    matches = np.zeros((cap, 2), np.int64)
    nm = c_int32(0)
    _check(lib().orc_pq_compare(
        _p(p), _p(t), c_long(p.size), c_int(num_categories), c_int64(ignored_label),
        c_int64(max_instances_per_category), c_int64(offset), c_int64(void_segment_id),
        _p(iou), _p(tp), _p(fn), _p(fp), _p(matches), c_int(cap), ctypes.byref(nm)))
    return iou, tp, fn, fp, {(int(g), int(q)) for g, q in matches[:nm.value]}


def pq_results(iou, tp, fn, fp, is_thing, ignored_label, suffix=''):
    """pq.py:182-187, 310-361 in float64 numpy."""
    iou, tp, fn, fp = (np.asarray(a, np.float64) for a in (iou, tp, fn, fp))
    is_thing = np.asarray(is_thing, bool)

    def realdiv(x, y):
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.where(np.abs(y) < 1e-10, 0.0, x / np.where(y == 0, 1.0, y))
    sq = realdiv(iou, tp)
    rq = realdiv(tp, tp + 0.5 * fn + 0.5 * fp)
    res = {'sq_per_class': sq, 'rq_per_class': rq, 'pq_per_class': sq * rq}
    valid = (tp + fn + fp) != 0
    valid_gt = (tp + fn) != 0
    if 0 <= ignored_label < len(tp):
        valid[ignored_label] = False
        valid_gt[ignored_label] = False
    sets = {f'all{suffix}': valid, f'things{suffix}': valid & is_thing,
            f'stuff{suffix}': valid & ~is_thing, f'all_with_gt{suffix}': valid_gt,
            f'things_with_gt{suffix}': valid_gt & is_thing,
            f'stuff_with_gt{suffix}': valid_gt & ~is_thing}
    for name, sel in sets.items():
        if sel.any():
            res[f'{name}_pq'] = res['pq_per_class'][sel].mean()
            res[f'{name}_sq'] = sq[sel].mean()
            res[f'{name}_rq'] = rq[sel].mean()
            res[f'{name}_num_categories'] = int(sel.sum())
        else:
            res[f'{name}_pq'] = res[f'{name}_sq'] = res[f'{name}_rq'] = 0
            res[f'{name}_num_categories'] = 0
    return res


def panoptic_postprocess(logits, heat, offset, orientation, is_thing, has_orientation,
                         threshold=0.1, nms_kernel_size=3, top_k=64,
                         apply_foreground_mask=False, normalized_offset=True,
                         offset_distance_threshold=None, max_instances_per_category=1 << 16,
                         cap=255):
    """Whole post-processing of a batch (panoptic.py:77-316, dense outputs + tables)."""
    logits = _c(logits, np.float32)
    heat = _c(heat, np.float32)
    offset = _c(offset, np.float32)
    ori = None if orientation is None else _c(orientation, np.float32)
    B, C, H, W = logits.shape
    thing = _c(is_thing, np.uint8)
    orient = _c(has_orientation, np.uint8) if has_orientation is not None else np.zeros((C,), np.uint8)
    assert thing.shape == (C,) and orient.shape == (C,)
    sem = np.empty((B, H, W), np.uint8)
    inst = np.empty((B, H, W), np.uint8)
    pan = np.empty((B, H, W), np.int64)
    centers = np.zeros((B, cap, 2), np.int32)
    n = np.zeros((B,), np.int32)
    area = np.zeros((B, cap + 1), np.int32)
    score = np.zeros((B, cap), np.float32)
    pairs = np.zeros((B, 256, 2), np.int64)
    n_pairs = np.zeros((B,), np.int32)
    present = np.zeros((B, 256), np.uint8)
    angle = np.zeros((B, 256), np.float32)
    use_thr = offset_distance_threshold is not None
    _check(lib().orc_panoptic_postprocess(
        _p(logits), _p(heat), _p(offset), _p(ori), c_int(B), c_int(C), c_int(H), c_int(W),
        _p(thing), _p(orient), c_float(threshold), c_int(nms_kernel_size), c_int(top_k),
        c_int(int(apply_foreground_mask)), c_int(int(normalized_offset)), c_int(int(use_thr)),
        c_float(offset_distance_threshold if use_thr else 0.0),
        c_int64(max_instances_per_category), _p(sem), _p(inst), _p(pan), _p(centers),
        c_int(cap), _p(n), _p(area), _p(score), _p(pairs), _p(n_pairs),
        _p(present) if ori is not None else None, _p(angle) if ori is not None else None))
    meta, ids, orientations = [], [], []
    for b in range(B):
        meta.append({i + 1: {'center_yx': (int(centers[b, i, 0]), int(centers[b, i, 1])),
                             'area': int(area[b, i + 1]),
                             'score': float(score[b, i])} for i in range(n[b])})
        ids.append({int(pairs[b, i, 0]): int(pairs[b, i, 1]) for i in range(n_pairs[b])})
        if ori is not None:
            orientations.append({i: float(angle[b, i]) for i in range(1, 256) if present[b, i]})
    return {'semantic_idx': sem, 'instance_idx': inst, 'panoptic': pan, 'ids': ids,
            'meta': meta, 'orientations': orientations if ori is not None else None}

# -*- coding: utf-8 -*-
"""Batch-dict conventions the post-processing relies on (host-side bookkeeping only).

Mirrors the behaviour of the reference helpers in `data/preprocessing/resize.py:22-78`:
entries at dataset resolution carry the suffix `_fullres`; the applied `Resize`
pre-processing step records the valid (un-padded) region of the network input.
"""
from typing import Any, Dict, Tuple

FULLRES_SUFFIX = '_fullres'
APPLIED_PREPROCESSING_KEY = '_applied_preprocessing'


def fullres_key(key: str) -> str:
    return key + FULLRES_SUFFIX


def fullres_shape(batch: Dict[str, Any], key: str) -> Tuple[int, int]:
    """(h, w) of the dataset resolution: taken from `<key>_fullres`, else from the rgb /
    depth full-res entries; ValueError if none is present (resize.py:30-47)."""
    for k in (key, 'rgb', 'depth'):
        entry = batch.get(fullres_key(k), None)
        if entry is not None:
            return tuple(entry.shape[-2:])
    raise ValueError(f"Unable to get fullres shape for `{key}`.")


def valid_region_slices(batch: Dict[str, Any]) -> Tuple[slice, slice]:
    """valid-region slices recorded by the Resize step of the first sample
    (resize.py:50-71); ValueError if the batch was never resized."""
    applied = batch.get(APPLIED_PREPROCESSING_KEY)
    if applied:
        for step in applied[0]:
            if step.get('type') == 'Resize':
                return step['valid_region_slice_y'], step['valid_region_slice_x']
    raise ValueError("Unable to get get valid region slices.")


def valid_region_and_fullres_shape(batch: Dict[str, Any], key: str):
    return valid_region_slices(batch), fullres_shape(batch, key)

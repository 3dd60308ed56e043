# -*- coding: utf-8 -*-
from .miou import MeanIntersectionOverUnion
from .pq import PanopticQuality, compare_and_accumulate
from .mae import MeanAbsoluteAngularError, PanopticQualityWithOrientationMAE
from .fused import PanopticEvaluation

__all__ = ['MeanIntersectionOverUnion', 'PanopticQuality', 'compare_and_accumulate',
           'MeanAbsoluteAngularError', 'PanopticQualityWithOrientationMAE', 'PanopticEvaluation']

# -*- coding: utf-8 -*-
"""Evaluation side of the reference's task helpers (task_helper/panoptic.py,
task_helper/instance.py, task_helper/semantic.py): the callers one level above the metric kernels.  Losses and
visualisation examples are outside the accelerated path and are not part of this package."""
from .base import TaskHelperBase
from .instance import InstanceTaskHelper
from .panoptic import PanopticTaskHelper
from .semantic import SemanticTaskHelper

__all__ = ['TaskHelperBase', 'InstanceTaskHelper', 'PanopticTaskHelper', 'SemanticTaskHelper']

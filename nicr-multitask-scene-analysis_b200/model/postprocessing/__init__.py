# -*- coding: utf-8 -*-
"""Post-processing classes of the panoptic path and their factory
(reference: model/postprocessing/__init__.py:24-44).  The other tasks' post-processing
(normal / scene / dense visual embedding) is outside this path and not provided."""
from typing import Any, Type

from ...utils.misc import partial_class
from .instance import InstancePostprocessing
from .panoptic import PanopticPostprocessing
from .semantic import SemanticPostprocessing

_REGISTRY = {
    'semantic': SemanticPostprocessing,
    'instance': InstancePostprocessing,
    'panoptic': PanopticPostprocessing,
}
_OUT_OF_SCOPE = ('dense-visual-embedding', 'normal', 'scene')


def get_postprocessing_class(name: str, **kwargs: Any) -> Type:
    """Class for task `name` with `kwargs` bound to its constructor."""
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"postprocessing '{name}' is outside the panoptic hot path this package replaces; "
            'use the reference implementation for it')
    if name not in _REGISTRY:
        raise ValueError(f"Unknown postprocessing: '{name}'")
    return partial_class(_REGISTRY[name], **kwargs)


__all__ = ['get_postprocessing_class', 'InstancePostprocessing', 'PanopticPostprocessing',
           'SemanticPostprocessing']

# -*- coding: utf-8 -*-
from functools import lru_cache


@lru_cache(maxsize=None)
def partial_class(cls, **bound):
    """Subclass of `cls` whose constructor has `bound` keyword arguments pre-applied
    (what `get_postprocessing_class` hands to the decoders, which instantiate it without
    arguments; reference: utils/_misc.py:11-21)."""
    if not bound:
        return cls

    def __init__(self, *args, **kwargs):
        cls.__init__(self, *args, **{**bound, **kwargs})

    return type(cls.__name__, (cls,), {'__init__': __init__, '__module__': cls.__module__})

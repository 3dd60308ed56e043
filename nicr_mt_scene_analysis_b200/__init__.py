"""Importable alias of the product package.

The package lives in the directory `nicr-multitask-scene-analysis_b200/` (the name the
project layout prescribes); a hyphen is not a legal Python identifier, so this shim loads
that directory under the importable name `nicr_mt_scene_analysis_b200` (the reference's
import name `nicr_mt_scene_analysis` + `_b200`).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_IMPL_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          'nicr-multitask-scene-analysis_b200')
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_IMPL_DIR, '__init__.py'),
                                     submodule_search_locations=[_IMPL_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

"""Importable alias of the package directory ``audio-few-shot-learning_b200/``.

The repository layout requires that directory name, which is not a valid Python
identifier; this shim loads it under the module name ``afsl_b200`` so that
``import afsl_b200.loops.loss`` etc. work with a single module identity.
"""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-few-shot-learning_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_root, "__init__.py"),
                                               submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

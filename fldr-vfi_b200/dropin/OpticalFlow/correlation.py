"""Drop-in for the reference's ``OpticalFlow.correlation`` module.

``OpticalFlow`` is a namespace package in the reference (no ``__init__.py``), so with this ``dropin/``
directory ahead of the reference checkout on ``sys.path`` the namespace merges both directories and
``OpticalFlow/PWCNet.py:4`` (``from . import correlation``) resolves here while ``PWCNet.py`` itself still
comes, untouched, from the reference.  Unlike the original, importing this module does not touch CUDA.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from fldr_vfi_b200.correlation import FunctionCorrelation, ModuleCorrelation, _FunctionCorrelation  # noqa: E402,F401

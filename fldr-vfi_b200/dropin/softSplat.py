"""Drop-in for the reference's top-level ``softSplat`` module.

Put this directory ahead of the reference checkout on ``sys.path`` and ``fLDRnet.py:22`` /
``utils.py:26`` (``from softSplat import Softsplat``) pick up the sm_100a implementation unmodified.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from fldr_vfi_b200.softSplat import Softsplat, FunctionSoftsplat, _FunctionSoftsplat  # noqa: E402,F401

"""Load the UNMODIFIED reference op modules (staged in baseline/_ref by fetch_ref.py) on a GPU box, with CuPy replaced
by baseline/cupy_shim.  Returns module objects under private names so they never shadow the drop-ins."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
_cache = {}


def available():
    return os.path.exists(os.path.join(REFDIR, "softSplat.py")) and os.path.exists(
        os.path.join(REFDIR, "OpticalFlow", "correlation.py"))


def _load(name, path):
    if name in _cache:
        return _cache[name]
    shim = os.path.join(HERE, "cupy_shim")
    had_real = "cupy" in sys.modules
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # the reference's regex literals are not raw strings
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    del had_real
    return mod


def softsplat_module():
    """The reference's softSplat.py: Softsplat / FunctionSoftsplat / _FunctionSoftsplat with its own CuPy kernels."""
    return _load("_fldr_reference_softSplat", os.path.join(REFDIR, "softSplat.py"))


def correlation_module():
    """The reference's OpticalFlow/correlation.py (touches CUDA at import, correlation.py:7-8: needs a GPU)."""
    return _load("_fldr_reference_correlation", os.path.join(REFDIR, "OpticalFlow", "correlation.py"))

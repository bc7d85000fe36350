"""ctypes binding of the C-ABI declared in ``include/fldr_b200.h``.

This is the same binding any foreign-language consumer would write; torch is used only
for device memory (``data_ptr``), strides and the current stream handle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLDR_B200_LIB") or os.path.join(_HERE, "libfldr_b200.so")     # override: A/B builds of the same ABI

c_float_p = ctypes.c_void_p
c_i64_p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes): every symbol include/fldr_b200.h declares
SYMBOLS = {
    "fldr_abi_version": (ctypes.c_int, []),
    "fldr_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "fldr_last_cuda_error": (ctypes.c_int, []),
    "fldr_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "fldr_get_option": (ctypes.c_int, [ctypes.c_char_p]),
    "fldr_splat_fwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "fldr_splat_fwd": (ctypes.c_int, [ctypes.c_int, c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_i64_p,
                                      c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fldr_splat_set_nonfinite_flag": (ctypes.c_int, [ctypes.c_void_p]),
    "fldr_splat_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "fldr_splat_bwd": (ctypes.c_int, [ctypes.c_int, c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_i64_p,
                                      c_float_p, c_float_p, c_float_p, c_i64_p, c_float_p, c_float_p, c_float_p,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fldr_corr81_fwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 4),
    "fldr_corr81_fwd": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fldr_corr81_fwd_act": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, ctypes.c_int64, ctypes.c_float,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "fldr_corr81_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 4),
    "fldr_corr81_bwd": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_i64_p,
                                       c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fldr_bwarp_fwd": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p]),
    "fldr_bwarp_bwd": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_float_p,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p]),
    "fldr_warp_metric_fwd": (ctypes.c_int, [c_float_p, c_i64_p, c_float_p, c_i64_p, c_float_p, c_i64_p, ctypes.c_float,
                                            c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p]),
    "fldr_occ_blend_fwd": (ctypes.c_int, [c_float_p, c_i64_p, ctypes.POINTER(ctypes.c_void_p), c_i64_p, c_float_p,
                                          ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "fldr_pca_features_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "fldr_pca_features_fwd": (ctypes.c_int, [c_float_p, c_i64_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fldr_bicubic_pyramid_fwd": (ctypes.c_int, [c_float_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
}

SPLAT_MODES = {"summation": 0, "average": 1, "linear": 2, "softmax": 3, "raw": 4}

_lib = None


class FldrError(RuntimeError):
    pass


def lib():
    """Load libfldr_b200.so.  No fallback: a missing library is an error, never a silent CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FldrError(f"{LIB_PATH} not found - build it with `python fldr-vfi_b200/build.py` "
                            "(there is no CPU / eager fallback for these ops)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.fldr_abi_version() != 1:
            raise FldrError("libfldr_b200.so ABI version mismatch")
        _lib = handle
    return _lib


_ext = None
_ext_tried = False


def ext():
    """The thin torch C++ extension over the same C ABI (``_fldr_torch_ext*.so``, built by build_ext.py), or None when it
    was not built - the ctypes binding above then serves every call (same library, same kernels, a few microseconds more
    host time per call).  FLDR_B200_NO_EXT=1 forces the ctypes binding."""
    global _ext, _ext_tried
    if not _ext_tried:
        _ext_tried = True
        if os.environ.get("FLDR_B200_NO_EXT", "0") in ("", "0"):
            import glob
            import importlib.util
            found = glob.glob(os.path.join(_HERE, "_fldr_torch_ext*.so"))
            if found:
                try:
                    lib()                                   # libfldr_b200.so first: the extension links against it
                    import torch  # noqa: F401  (libtorch / libcudart must be loaded before the extension)
                    spec = importlib.util.spec_from_file_location("_fldr_torch_ext", found[0])
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    if mod.abi_version() == 1:
                        _ext = mod
                except Exception:                           # a stale / incompatible build: the ctypes binding serves the calls
                    _ext = None
    return _ext


def check(status):
    if status != 0:
        l = lib()
        msg = l.fldr_status_string(status).decode()
        raise FldrError(f"{msg} (status {status}, cudaError {l.fldr_last_cuda_error()})")


def strides(t):
    return (ctypes.c_int64 * 4)(*t.stride())


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())

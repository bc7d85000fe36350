"""Minimal stand-in for the parts of CuPy the reference's two op files use, so that the UNMODIFIED reference files
(softSplat.py, OpticalFlow/correlation.py) import and run on a box without CuPy (none in this image, no network).

Same CUDA C source, same launch geometry: the kernel strings still go through the reference's own ``cupy_kernel``
templating and are compiled by NVRTC (as CuPy would do) via ``torch.cuda._compile_kernel``; launches go through
``cuLaunchKernel`` on the stream handle the reference passes.  Used only by the baseline harness and the
ours-vs-reference GPU tests - never by the product path.

Covered API (call sites in the reference): ``cupy.memoize(for_each_device=True)`` (softSplat.py:215,
correlation.py:287), ``cupy.RawModule(code=...).get_function(name)`` (softSplat.py:217),
``cupy.cuda.compile_with_cache(code).get_function(name)`` (correlation.py:289), ``cupy.int32`` (softSplat.py:245).
"""
import ctypes
import functools


class int32(int):
    """Marks a kernel argument as a 32-bit integer (plain Python ints are device pointers in the reference's calls)."""


def memoize(for_each_device=False):
    def deco(fn):
        cache = {}

        @functools.wraps(fn)
        def wrapper(*args):
            import torch
            key = (torch.cuda.current_device() if for_each_device else 0,) + args
            if key not in cache:
                cache[key] = fn(*args)
            return cache[key]
        return wrapper
    return deco


class _Function:
    def __init__(self, kernel):
        self._kernel = kernel       # torch.cuda._utils._CudaKernel

    def __call__(self, grid, block, args, shared_mem=0, stream=None):
        import torch
        libcuda = torch.cuda._utils._get_gpu_runtime_library()
        holders, ptrs = [], []
        for a in args:
            if isinstance(a, int32):
                h = ctypes.c_int(int(a))
            elif a is None:
                h = ctypes.c_void_p(0)
            elif isinstance(a, int):
                # correlation.py passes n / intSample as plain ints and pointers as plain ints too; CuPy maps Python
                # ints to 64-bit, which is what a pointer needs and is harmless for the little-endian `const int n`
                h = ctypes.c_longlong(a)
            else:
                raise TypeError(f"unsupported kernel argument {type(a)}")
            holders.append(h)
            ptrs.append(ctypes.cast(ctypes.byref(h), ctypes.c_void_p))
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        sptr = getattr(stream, "ptr", None)
        if sptr is None:
            sptr = torch.cuda.current_stream().cuda_stream
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        err = libcuda.cuLaunchKernel(self._kernel.func, grid[0], grid[1], grid[2], block[0], block[1], block[2],
                                     int(shared_mem), ctypes.c_void_p(sptr), arr, None)
        if err != 0:
            raise RuntimeError(f"cuLaunchKernel failed with {err}")


_PREAMBLE = "#include <assert.h>\n"


class RawModule:
    def __init__(self, code, options=(), name_expressions=None, **kw):
        self._code = code

    def get_function(self, name):
        import torch
        try:
            k = torch.cuda._compile_kernel(_PREAMBLE + self._code, name, cuda_include_dirs=["/usr/local/cuda/include"])
        except Exception:
            # no header path: keep the kernels running with asserts compiled out (noted in the baseline report)
            k = torch.cuda._compile_kernel("#define assert(x) ((void)0)\n" + self._code, name)
        return _Function(k)


class _Cuda:
    @staticmethod
    def compile_with_cache(code, options=(), **kw):
        return RawModule(code)


cuda = _Cuda()


# ---- non-hot-path helpers the reference model files reference at module / setup time (fLDRnet.py:232-264,
# pca_comp.py:353,372; useful.py:68-73).  numpy-backed; never used inside the timed forward.
def asnumpy(a):
    import numpy as np
    return np.asarray(a)


def asarray(a):
    import numpy as np
    return np.asarray(a)


class _Pool:
    def free_all_blocks(self):
        pass


def get_default_memory_pool():
    return _Pool()

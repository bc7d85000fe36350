"""fldr-vfi_b200: B200-native (sm_100a) replacement for fLDR-VFI's two CuPy-JIT operators.

Only the hot path lives here:

  csrc/            hand-written CUDA kernels + the C-ABI of include/fldr_b200.h  -> libfldr_b200.so
  _lib.py          ctypes binding of that C-ABI (fails loudly when the library is missing)
  softSplat.py     host mirror of the reference's softSplat.py  (Softsplat / FunctionSoftsplat / _FunctionSoftsplat)
  correlation.py   host mirror of OpticalFlow/correlation.py    (ModuleCorrelation / FunctionCorrelation / _FunctionCorrelation)
  dropin/          ``softSplat`` and ``OpticalFlow.correlation`` modules under the reference's own import names
  sharding.py      frame-pair partitioning across ranks (no collective on the data path)

There is no CPU fallback: every op raises if ``libfldr_b200.so`` is absent or the tensors are not CUDA fp32.
"""
__version__ = "0.1.0"

"""Build ``libfldr_b200.so`` (the C-ABI library of include/fldr_b200.h) in-tree with nvcc for sm_100a.

    python fldr-vfi_b200/build.py [--force]

The library links the CUDA runtime statically and has no torch / Python dependency: the same file
serves the ctypes host layer here and any cgo / JNI / N-API consumer.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfldr_b200.so")
STAMP = os.path.join(HERE, ".libfldr_b200.stamp")
SOURCES = ["cabi.cu", "splat.cu", "corr.cu", "warp.cu", "blend.cu", "pca.cu", "pyramid.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-Xlinker", "-soname=libfldr_b200.so", "--use_fast_math=false"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]  # precise math: parity bar is 1e-5 relative


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/fldr_b200.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print("[fldr_b200/build]", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    open(os.path.join(HERE, "build.log"), "w").write(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libfldr_b200.so")
    open(STAMP, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

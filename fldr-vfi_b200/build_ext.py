"""Build ``_fldr_torch_ext`` - the thin PyTorch C++ extension over ``libfldr_b200.so`` - in-tree (host C++ only, no CUDA
sources: it links the C-ABI library next to it through an $ORIGIN rpath).

    python fldr-vfi_b200/build_ext.py [--force]
"""
import glob
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_ext", "fldr_torch_ext.cpp")
NAME = "_fldr_torch_ext"
STAMP = os.path.join(HERE, "." + NAME + ".stamp")


def target():
    found = glob.glob(os.path.join(HERE, NAME + "*.so"))
    return found[0] if found else None


def _digest():
    import torch
    h = hashlib.sha256()
    for p in (SRC, os.path.join(HERE, "..", "include", "fldr_b200.h")):
        h.update(open(p, "rb").read())
    h.update(torch.__version__.encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    dig = _digest()
    if not force and target() and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return target()
    from torch.utils import cpp_extension
    build_dir = os.path.join(HERE, "build_ext_tmp")
    os.makedirs(build_dir, exist_ok=True)
    try:
        cpp_extension.load(
            name=NAME, sources=[SRC], build_directory=build_dir, verbose=verbose, with_cuda=True,
            extra_cflags=["-O2", "-std=c++17"],
            extra_ldflags=["-L" + HERE, "-lfldr_b200", "-Wl,-rpath,'$$ORIGIN'"],      # $$: ninja escape
            is_python_module=False)
    except OSError:
        pass      # load() also tries to dlopen the result from the scratch directory, where $ORIGIN does not see the C-ABI library
    built = os.path.join(build_dir, NAME + ".so")
    dst = os.path.join(HERE, NAME + ".so")
    shutil.copyfile(built, dst)
    shutil.rmtree(build_dir, ignore_errors=True)
    open(STAMP, "w").write(dig)
    return dst


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Time fldr_corr81_fwd of a given build of the library on the bench's two large pyramid levels.

    python tools/corr_probe.py build/libfldr_ck4_s6.so [more.so ...]

Each library is loaded in a fresh subprocess (options and kernels are per-process).  Results are checked against the
first library's output bit-for-bit where the summation order is the same, else to 1e-5 relative.
"""
import ctypes
import os
import subprocess
import sys

LEVELS = [("C32", 2, 32, 576, 1024), ("C64", 2, 64, 288, 512), ("C96", 2, 96, 144, 256)]


def worker(path):
    import torch
    lib = ctypes.CDLL(os.path.abspath(path))
    i64x4 = ctypes.c_int64 * 4
    out_line = [os.path.basename(path)]
    for name, B, C, H, W in LEVELS:
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(B, C, H, W, device="cuda", generator=g)
        b = torch.randn(B, C, H, W, device="cuda", generator=g)
        o = torch.empty(B, 81, H, W, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def call():
            r = lib.fldr_corr81_fwd(ctypes.c_void_p(a.data_ptr()), i64x4(*a.stride()), ctypes.c_void_p(b.data_ptr()),
                                    i64x4(*b.stride()), ctypes.c_void_p(o.data_ptr()), B, C, H, W, None,
                                    ctypes.c_size_t(0), ctypes.c_void_p(st))
            assert r == 0, r
        for _ in range(3):
            call()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        ref = torch.einsum("bchw,bchw->bhw", a[:, :, 100:110, 200:210], b[:, :, 101:111, 198:208]) / C   # dy=+1, dx=-2
        got = o[:, (1 + 4) * 9 + (-2 + 4), 100:110, 200:210]
        err = float((got - ref).abs().max())
        out_line.append(f"{name} {ts[len(ts) // 2]:7.1f} us (min {ts[0]:.1f}, err {err:.1e}, sum {float(o.double().sum()):.6f})")
    print("  ".join(out_line), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        for p in sys.argv[1:]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", p], check=False)

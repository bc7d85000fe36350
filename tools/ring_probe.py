"""GPU probe of the streaming (ring) splat kernel: parity on small shapes, flag word, timings against the whole-frame
path for a sweep of the tuning options.  Run on the B200 box:  python tools/ring_probe.py [--quick]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fldr_vfi_b200._lib as L  # noqa: E402
import fldr_vfi_b200.softSplat as S  # noqa: E402
from oracle import splat_oracle as so, synth  # noqa: E402

lib = L.lib()


def plan(mode, N, C, H, W, hm):
    info = (ctypes.c_int64 * 8)()
    assert lib.fldr_splat_fwd_plan(mode, N, C, H, W, int(hm), info) == 0
    return list(info)


def raw_call(x, fl, z, mode=3):
    """fldr_splat_fwd through the C ABI with our own workspace; returns (out, flag word)."""
    N, C, H, W = x.shape
    info = plan(mode, N, C, H, W, z is not None)
    ws = torch.zeros(info[6], dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x)
    st = lib.fldr_splat_fwd(mode, L.ptr(x), L.strides(x), L.ptr(fl), L.strides(fl), L.ptr(z), None if z is None else L.strides(z),
                            L.ptr(out), None, N, C, H, W, L.ptr(ws), info[6], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(st)
    torch.cuda.synchronize()
    flag = int(ws[info[3]:info[3] + 4].view(torch.int32)[0]) if info[0] == 2 else -1
    return out, flag, info


def close(a, b, mag=1.0):
    a, b = a.double().cpu(), b.double().cpu()
    err = (a - b).abs()
    bad = err > 1e-4 * mag + 1e-5 * b.abs()
    return int(bad.sum()), float(err.max())


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    quick = "--quick" in sys.argv
    torch.cuda.set_device(0)
    lib.fldr_set_option(b"splat_fused_max", 0)      # force the ring path on small frames too
    print("== parity on small shapes (ring path forced)")
    cases = [((1, 3, 256, 448), "F1", True), ((1, 16, 256, 448), "F1", False), ((1, 48, 32, 56), "F2", False),
             ((2, 3, 64, 96), "F3", True), ((2, 5, 36, 48), "FB", False), ((3, 7, 8, 132), "F2", False),
             ((2, 3, 40, 260), "FB", True), ((1, 3, 4, 4), "F0", True), ((2, 48, 72, 128), "F1", False)]
    for shape, regime, hm in cases:
        N, C, H, W = shape
        x = synth.features(N, C, H, W, seed=11)
        fl = synth.flow(N, H, W, regime, seed=12)
        z = synth.metric(N, H, W, seed=13) if hm else None
        ref = so.function_softsplat(x, fl, z, "softmax")
        out, flag, info = raw_call(x.cuda(), fl.cuda(), None if z is None else z.cuda())
        nbad, emax = close(out, ref)
        print(f"  {shape} {regime} metric={hm}: path={info[0]} reach={info[1]} flag={flag} bad={nbad} maxerr={emax:.2e}", flush=True)
    for mode, name in ((0, "summation"), (1, "average"), (2, "linear"), (4, "raw")):
        N, C, H, W = 2, 6, 40, 64
        x = synth.features(N, C, H, W, seed=21)
        fl = synth.flow(N, H, W, "F2", seed=22)
        z = synth.metric(N, H, W, seed=23) if mode == 2 else None
        if mode == 2:
            x = x[:, :3].contiguous()
        ref = so.function_softsplat(x, fl, z, name) if mode != 4 else so.splat_raw(x, fl)
        out, flag, info = raw_call(x.cuda(), fl.cuda(), None if z is None else z.cuda(), mode)
        nbad, emax = close(out, ref, mag=max(1.0, float(ref.abs().max())))
        print(f"  mode {name}: path={info[0]} flag={flag} bad={nbad} maxerr={emax:.2e}", flush=True)
    lib.fldr_set_option(b"splat_fused_max", 40000)

    print("== bounded ring: tall frames")
    for (N, H, W, bump) in ((1, 1152, 4096, 0.0), (1, 1152, 4096, 260.0), (3, 640, 4096, 0.0)):
        x = synth.image(N, 3, H, W, seed=81)
        z = synth.metric(N, H, W, seed=82)
        fl = synth.flow(N, H, W, "F1", seed=83)
        if bump:
            fl[:, 1, 300:340, 1000:1400] += bump
            fl[:, 1, 900:930, 2000:2100] -= 400.0
        ref = so.function_softsplat(x, fl, z, "softmax")
        out, flag, info = raw_call(x.cuda(), fl.cuda(), z.cuda())
        nbad, emax = close(out, ref)
        print(f"  N={N} {H}x{W} bump={bump}: path={info[0]} reach={info[1]} flag={flag} bad={nbad} maxerr={emax:.2e}", flush=True)

    print("== 4K image splat (C=3 + metric, 2304x4096)")
    H, W = 2304, 4096
    x = synth.image(1, 3, H, W, seed=56).cuda()
    z = synth.metric(1, H, W, seed=58).cuda()
    alg = 4 * H * W * 9
    flows = {"F1": synth.flow(1, H, W, "F1", seed=57).cuda(), "F2": synth.flow(1, H, W, "F2", seed=59).cuda()}
    if not quick:
        flows["F0"] = torch.zeros(1, 2, H, W, device="cuda")
    lib.fldr_set_option(b"splat_stream", 0)
    refs = {}
    for k, fl in flows.items():
        refs[k] = S.FunctionSoftsplat(x, fl, z, "softmax")
        med, mn = timeit(lambda: S.FunctionSoftsplat(x, fl, z, "softmax"))
        print(f"  whole-frame {k}: {med:.1f} us median, {mn:.1f} min -> {alg / med / 1e3:.0f} GB/s ({alg / med / 1e3 / 6549.1:.3f})", flush=True)
    lib.fldr_set_option(b"splat_stream", 1)
    sweeps = [(0, 0)] if quick else [(0, 0), (0, 8), (0, 10), (0, 18), (0, 24), (68, 0), (68, 14), (18, 0)]
    for ring_mb, lag in sweeps:
        lib.fldr_set_option(b"splat_ring_mb", ring_mb)
        lib.fldr_set_option(b"splat_lag", lag)
        for k, fl in flows.items():
            out, flag, info = raw_call(x, fl, z)
            nbad, emax = close(out, refs[k])
            med, mn = timeit(lambda: S.FunctionSoftsplat(x, fl, z, "softmax"))
            print(f"  ring mb={ring_mb} lag={lag} {k}: reach={info[1]} rows={info[4]} flag={flag} bad={nbad} maxerr={emax:.1e} | "
                  f"{med:.1f} us median, {mn:.1f} min -> {alg / med / 1e3:.0f} GB/s ({alg / med / 1e3 / 6549.1:.3f})", flush=True)
    lib.fldr_set_option(b"splat_ring_mb", 0)
    lib.fldr_set_option(b"splat_lag", 0)

    print("== feature splats (C=48, no metric) and training shapes")
    shapes = [(1, 48, 288, 512, False), (1, 48, 144, 256, False), (1, 48, 72, 128, False), (32, 3, 512, 512, True),
              (32, 3, 256, 256, True), (32, 48, 64, 64, False)]
    for (N, C, h, w, hm) in shapes:
        xf = (synth.features(N, C, h, w, seed=71) if C != 3 else synth.image(N, C, h, w, seed=71)).cuda()
        ff = (synth.flow(N, h, w, "F1", seed=72) * 8).cuda()
        zf = synth.metric(N, h, w, seed=73).cuda() if hm else None
        algf = 4 * N * h * w * (2 * C + 2 + (1 if hm else 0))
        res = {}
        for stream in (0, 1):
            lib.fldr_set_option(b"splat_stream", stream)
            y = S.FunctionSoftsplat(xf, ff, zf, "softmax")
            med, mn = timeit(lambda: S.FunctionSoftsplat(xf, ff, zf, "softmax"))
            res[stream] = (y, med, mn)
        nbad, emax = close(res[1][0], res[0][0])
        info = plan(3, N, C, h, w, hm)
        print(f"  {(N, C, h, w)}: whole {res[0][1]:.1f} us | ring(path={info[0]} reach={info[1]}) {res[1][1]:.1f} us ({res[1][2]:.1f} min) "
              f"-> {algf / res[1][1] / 1e3:.0f} GB/s ({algf / res[1][1] / 1e3 / 6549.1:.3f}); bad={nbad} maxerr={emax:.1e}", flush=True)
    lib.fldr_set_option(b"splat_stream", 1)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"done in {time.time() - t0:.1f} s")

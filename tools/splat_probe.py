"""GPU probe of the splat forward: per-shape timings of the TMA tile scatter path vs the plain-load path.  python tools/splat_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
import fldr_vfi_b200.softSplat as S
from oracle import synth
lib = L.lib()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


shapes = [(1, 3, 2304, 4096, True, "F1"), (1, 3, 2304, 4096, True, "F2"), (1, 3, 2304, 4096, True, "F0"), (1, 48, 288, 512, False, "F1"),
          (1, 48, 144, 256, False, "F1"), (1, 48, 72, 128, False, "F1"), (32, 3, 512, 512, True, "F1"), (32, 3, 256, 256, True, "F1"),
          (32, 48, 64, 64, False, "F1"), (1, 48, 36, 64, False, "F1"), (1, 48, 18, 32, False, "F1"), (32, 3, 64, 64, True, "F1"),
          (32, 3, 128, 128, True, "F1")]
if len(sys.argv) > 1 and sys.argv[1] == "4k":
    shapes = [s for s in shapes if s[2] == 2304 and s[5] != "F0"] + [(32, 3, 512, 512, True, "F1")]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    shapes = [s for s in shapes if s[2] * s[3] <= 25600]
for (N, C, h, w, hm, reg) in shapes:
    x = (synth.features(N, C, h, w, seed=71) if C != 3 else synth.image(N, C, h, w, seed=71)).cuda()
    f = synth.flow(N, h, w, reg, seed=72)
    if h < 2304 and reg == "F1": f = f * 8
    f = f.cuda()
    z = synth.metric(N, h, w, seed=73).cuda() if hm else None
    alg = 4 * N * h * w * (2 * C + 2 + (1 if hm else 0))
    res = {}
    for tma in (0, 1):
        lib.fldr_set_option(b"splat_tma", tma)
        for pf in ((0, -1) if tma else (0,)):
            lib.fldr_set_option(b"splat_pf_rows", pf)
            y = S.FunctionSoftsplat(x, f, z, "softmax")
            med, mn = timeit(lambda: S.FunctionSoftsplat(x, f, z, "softmax"))
            res[(tma, pf)] = (y, med, mn)
    lib.fldr_set_option(b"splat_pf_rows", 0)
    err = float((res[(1, 0)][0] - res[(0, 0)][0]).abs().max())
    print(f"{(N, C, h, w)} {reg}: plain {res[(0,0)][1]:.1f} us | tma {res[(1,0)][1]:.1f} us (min {res[(1,0)][2]:.1f}; no prefetch {res[(1,-1)][1]:.1f}) "
          f"-> {alg / res[(1,0)][1] / 1e3:.0f} GB/s ({alg / res[(1,0)][1] / 1e3 / 6549.1:.3f}); max|diff| {err:.1e}", flush=True)

"""GPU probe of the tile-owner splat path against the three-pass path.  python tools/owner_probe.py"""
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


modes = [int(a) for a in sys.argv[1:]] or [0, 2]
shapes = [(1, 3, 2304, 4096, True, "F1"), (1, 3, 2304, 4096, True, "F0"), (32, 3, 512, 512, True, "F1"), (1, 3, 2304, 4096, True, "F2")]
for (N, C, h, w, hm, reg) in shapes:
    x = synth.image(N, C, h, w, seed=71).cuda()
    f = synth.flow(N, h, w, reg, seed=72)
    if h < 2304 and reg == "F1": f = f * 8
    f = f.cuda()
    z = synth.metric(N, h, w, seed=73).cuda() if hm else None
    alg = 4 * N * h * w * (2 * C + 2 + (1 if hm else 0))
    res = {}
    for own in modes:
        lib.fldr_set_option(b"splat_owner", own)
        y = S.FunctionSoftsplat(x, f, z, "softmax")
        torch.cuda.synchronize()
        med, mn = timeit(lambda: S.FunctionSoftsplat(x, f, z, "softmax"))
        res[own] = (y, med, mn)
        print(f"{(N, C, h, w)} {reg} owner={own}: {med:.1f} us (min {mn:.1f}) -> {alg / med / 1e3:.0f} GB/s ({alg / med / 1e3 / 6549.1:.3f})", flush=True)
    lib.fldr_set_option(b"splat_owner", 0)
    if len(modes) > 1:
        err = float((res[modes[0]][0] - res[modes[-1]][0]).abs().max())
        print(f"    max|diff| between modes {err:.2e}", flush=True)

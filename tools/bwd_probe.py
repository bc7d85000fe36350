"""GPU probe of the splat backward (cfg5 image splat 32x3x512^2 with metric, all three gradients), op-level entry timed with CUDA
events (no autograd engine in the timed region).  python tools/bwd_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
import fldr_vfi_b200.softSplat as S
from oracle import synth


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


for (N, C, H, W, hm, need) in [(32, 3, 512, 512, True, (True, True, True)), (32, 3, 512, 512, False, (True, True, False)),
                               (32, 48, 64, 64, False, (True, False, False)), (1, 3, 2304, 4096, True, (True, True, True))]:
    x = (synth.image(N, C, H, W, seed=1) if C == 3 else synth.features(N, C, H, W, seed=1)).cuda()
    fl = (synth.flow(N, H, W, "F1", seed=2) * 4).cuda()
    z = synth.metric(N, H, W, seed=3).cuda() if hm else None
    g = synth.grad((N, C, H, W), seed=4).cuda()
    out, norm = S._splat_forward(3, x, fl, z, True)
    med, mn = timeit(lambda: S._splat_backward(3, x, fl, z, out, norm, g, need))
    nb = 4 * N * H * W * ((4 * C + 7) if need[1] else (3 * C + 3))
    print(f"{(N, C, H, W)} metric={hm} need={need}: {med:.1f} us (min {mn:.1f}) -> {nb / med / 1e3:.0f} GB/s ({nb / med / 1e3 / 6549.1:.3f})", flush=True)

"""GPU probe of the small / medium-frame splats: three-launch path vs cooperative single launch (splat_fused_max) (default threshold / forced); host-visible time per call and device time with the host running ahead."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
import fldr_vfi_b200.softSplat as S
from oracle import synth
lib = L.lib()


def t_call(fn, iters=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


def t_dev(fn, iters=10, reps=8):
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(600000)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return sorted(ts)[len(ts) // 2]


shapes = [(1, 48, 288, 512, False), (1, 48, 144, 256, False), (1, 48, 72, 128, False), (1, 48, 36, 64, False), (1, 48, 18, 32, False),
          (32, 48, 64, 64, False), (32, 48, 32, 32, False), (32, 48, 16, 16, False), (32, 3, 64, 64, True), (32, 3, 128, 128, True), (32, 3, 256, 256, True)]
for (N, C, h, w, hm) in shapes:
    x = (synth.features(N, C, h, w, seed=71) if C != 3 else synth.image(N, C, h, w, seed=71)).cuda()
    f = (synth.flow(N, h, w, "F1", seed=72) * 8).cuda()
    z = synth.metric(N, h, w, seed=73).cuda() if hm else None
    fn = lambda: S.FunctionSoftsplat(x, f, z, "softmax")
    row = []
    ref = None
    for name, fm in (("3-launch", 0), ("coop", 40000), ("coop-all", 1 << 30)):
        lib.fldr_set_option(b"splat_fused_max", fm)
        y = fn()
        if ref is None: ref = y
        row.append(f"{name}: {t_call(fn):.1f}/{t_dev(fn):.1f} (diff {float((y - ref).abs().max()):.1e})")
    lib.fldr_set_option(b"splat_fused_max", 40000)
    print(f"({N},{C},{h},{w}) per-call us / device us  " + " | ".join(row), flush=True)

"""cfg5 (train_it.py-shaped step): 32 x 512^2 crops, splat + correlation forward AND backward, ours vs the reference
kernels on the same GPU.  Prints per-call device times (CUDA events, median of 5 after warm-up)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.softSplat as S
import fldr_vfi_b200.correlation as C
from baseline import ref_gpu
from oracle import synth

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

def splat_case(mod, N, C, H, W, metric, need_flow):
    x = synth.image(N, C, H, W, seed=1).cuda().requires_grad_(True)
    fl = (synth.flow(N, H, W, "F1", seed=2) * 4).cuda().requires_grad_(need_flow)
    z = synth.metric(N, H, W, seed=3).cuda().requires_grad_(True) if metric else None
    g = synth.grad((N, C, H, W), seed=4).cuda()
    sp = mod.Softsplat()
    fwd = timed(lambda: sp(x, fl, z))
    def both():
        y = sp(x, fl, z)
        torch.autograd.grad(y, [t for t in (x, fl, z) if t is not None and t.requires_grad], g)
    return fwd, timed(both)

def corr_case(mod, B, C, H, W):
    a = synth.features(B, C, H, W, seed=3).cuda().requires_grad_(True)
    b = synth.features(B, C, H, W, seed=5).cuda().requires_grad_(True)
    g = synth.grad((B, 81, H, W), seed=4).cuda()
    fwd = timed(lambda: mod.FunctionCorrelation(tensorFirst=a, tensorSecond=b))
    def both():
        o = mod.FunctionCorrelation(tensorFirst=a, tensorSecond=b)
        torch.autograd.grad(o, [a, b], g)
    return fwd, timed(both)

mods = {"ours": (S, C)}
if ref_gpu.available(): mods["reference"] = (ref_gpu.softsplat_module(), ref_gpu.correlation_module())
rows = []
for name, (s, c) in mods.items():
    for (N, Cc, H, W, metric, nf, tag) in [(32, 3, 512, 512, True, True, "image splat 32x3x512^2 +z"), (32, 3, 256, 256, True, True, "image splat 32x3x256^2 +z"),
                                           (32, 48, 64, 64, False, False, "feature splat 32x48x64^2"), (32, 3, 512, 512, False, True, "endflow-loss splat 32x3x512^2")]:
        f, fb = splat_case(s, N, Cc, H, W, metric, nf)
        rows.append((name, tag, f, fb))
    for (B, Cc, H, W) in [(64, 32, 128, 128), (64, 64, 64, 64), (64, 96, 32, 32), (64, 196, 8, 8)]:
        f, fb = corr_case(c, B, Cc, H, W)
        rows.append((name, f"corr {B}x{Cc}x{H}x{W}", f, fb))
for r in rows: print(f"{r[0]:10s} {r[1]:34s} fwd {r[2]*1000:9.1f} us   fwd+bwd {r[3]*1000:9.1f} us")

# ---- next row: bwarp at the training crop shape (32 x 3 x 512^2), ours vs the reference's method AS WRITTEN
# (fLDRnet.py:546-581: mesh grid and ones tensor built on the CPU and copied to the device on every call) and vs the same
# operator sequence with those tensors already on the device
import fldr_vfi_b200.warp as Wp


from baseline import ref_src

N, Cc, H, W = 32, 3, 512, 512
x = synth.image(N, Cc, H, W, seed=1).cuda().requires_grad_(True)
fl = (synth.flow(N, H, W, "F1", seed=2) * 4).cuda().requires_grad_(True)
g = synth.grad((N, Cc, H, W), seed=4).cuda()
dev = x.device
ref_as_written, ref_on_device = ref_src.bwarp(dev), ref_src.bwarp(dev, create_on_device=True)
for name, fn in (("ours", lambda: Wp.bwarp(x, fl, True)), ("reference (as written)", lambda: ref_as_written(x, fl, True)),
                 ("reference text, grid/ones on device", lambda: ref_on_device(x, fl, True))):
    f = timed(fn)
    fb = timed(lambda: torch.autograd.grad(fn(), [x, fl], g))
    print(f"{name:36s} bwarp 32x3x512^2   fwd {f*1000:9.1f} us   fwd+bwd {fb*1000:9.1f} us")

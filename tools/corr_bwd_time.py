"""Device time of the correlation backward per gradient at the cfg5 shapes (CUDA events, host running ahead)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
from oracle import synth
import ctypes
lib = L.lib()          # FLDR_B200_LIB=... selects another build of the same ABI


def corr81_bwd(f1, f2, go, need_first, need_second):
    B, Cc, H, W = f1.shape
    g1 = torch.empty_like(f1) if need_first else None
    g2 = torch.empty_like(f1) if need_second else None
    nb = lib.fldr_corr81_bwd_workspace_bytes(B, Cc, H, W)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=f1.device)
    st = lib.fldr_corr81_bwd(L.ptr(f1), L.strides(f1), L.ptr(f2), L.strides(f2), L.ptr(go), L.strides(go), L.ptr(g1), L.ptr(g2),
                             B, Cc, H, W, L.ptr(ws), nb, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(st)
    return g1, g2


for (B, Cc, H, W) in [(64, 32, 128, 128), (64, 64, 64, 64), (2, 32, 576, 1024)]:
    f1 = synth.features(B, Cc, H, W, seed=3).cuda(); f2 = synth.features(B, Cc, H, W, seed=5).cuda()
    go = synth.grad((B, 81, H, W), seed=4).cuda()
    row = []
    for need in ((True, False), (False, True), (True, True)):
        fn = lambda: corr81_bwd(f1, f2, go, *need)
        fn(); ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)
            a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 500)
        row.append(f"{'first' if need[0] else ''}{'+' if all(need) else ''}{'second' if need[1] else ''} {sorted(ts)[5]:.1f} us")
    alg = 4 * B * H * W * (4 * Cc + 81)
    t_both = float(row[-1].split()[-2])
    print(f"({B},{Cc},{H},{W}): " + " | ".join(row) + f"  -> {alg / t_both / 1e3:.0f} GB/s ({alg / t_both / 1e3 / 6549.1:.3f}), {2 * 162.0 * Cc * B * H * W / t_both / 1e6:.1f} TFLOP/s")

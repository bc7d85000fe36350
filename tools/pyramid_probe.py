"""Timing of the device-side input pyramid at the padded 4K frame pair (tools; CUDA events, median of 20)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import fldr_vfi_b200.pyramid as Py

def med(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)          # the host runs ahead of the device: the events bracket device time, not launch overhead
        e0.record(); fn(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 4)
    return sorted(ts)[len(ts) // 2]

H, W = 2304, 4096
fr = torch.rand(1, 3, 2, H, W, device="cuda")
sc = [8, 16, 32, 64, 128, 256]
for n in (5, 3, 1):
    ms = med(lambda: Py.input_pyramid(fr, sc, n))
    nb = 4 * 6 * sum((H >> k) * (W >> k) for k in range(n + 1))
    print(f"pow2 n_levels={n}: {ms*1e3:.1f} us  {nb/ms/1e6:.0f} GB/s ({nb/ms/1e6/6549.1:.3f})")
pl = fr.permute(0, 2, 1, 3, 4).reshape(2, 3, H, W)
ms = med(lambda: [F.interpolate(pl, scale_factor=1.0 / (1 << k), mode="bicubic") for k in range(1, 6)])
print(f"torch cuda interpolate x5: {ms*1e3:.1f} us")
ms = med(lambda: Py.bicubic_levels(fr, [1 / 3]))
print(f"generic 1/3: {ms*1e3:.1f} us")

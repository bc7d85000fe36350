"""Host-side cost per op call (tiny inputs, so GPU time is negligible): ours vs the reference ops on the same GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.softSplat as S
import fldr_vfi_b200.correlation as C
from baseline import ref_gpu

def bench(fn, n=300, empty_cache=False):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if empty_cache: torch.cuda.empty_cache()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

x = torch.randn(1, 48, 18, 32, device="cuda"); fl = torch.randn(1, 2, 18, 32, device="cuda")
xi = torch.randn(1, 3, 64, 96, device="cuda"); fli = torch.randn(1, 2, 64, 96, device="cuda"); z = torch.randn(1, 1, 64, 96, device="cuda")
f1 = torch.randn(2, 32, 16, 16, device="cuda"); f2 = torch.randn(2, 32, 16, 16, device="cuda")
mods = {"ours": (S, C)}
if ref_gpu.available(): mods["reference"] = (ref_gpu.softsplat_module(), ref_gpu.correlation_module())
with torch.no_grad():
    for name, (s, c) in mods.items():
        sp = s.Softsplat()
        print(f"{name:10s} feature splat 18x32 C48   : {bench(lambda: sp(x, fl)):8.1f} us/call   with empty_cache: {bench(lambda: sp(x, fl), 100, True):8.1f}")
        print(f"{name:10s} image splat 64x96 C3+z    : {bench(lambda: sp(xi, fli, z)):8.1f} us/call   with empty_cache: {bench(lambda: sp(xi, fli, z), 100, True):8.1f}")
        print(f"{name:10s} correlation 2x32x16x16    : {bench(lambda: c.FunctionCorrelation(tensorFirst=f1, tensorSecond=f2)):8.1f} us/call")
xl = torch.randn(1, 3, 512, 768, device="cuda"); fll = torch.randn(1, 2, 512, 768, device="cuda"); zl = torch.randn(1, 1, 512, 768, device="cuda")
with torch.no_grad():
    for name, (s, c) in mods.items():
        sp = s.Softsplat()
        print(f"{name:10s} image splat 512x768       : {bench(lambda: sp(xl, fll, zl), 100):8.1f} us/call   with empty_cache: {bench(lambda: sp(xl, fll, zl), 50, True):8.1f}")

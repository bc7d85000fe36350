import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fldr_vfi_b200.softSplat as S
x = torch.randn(1, 48, 18, 32, device="cuda"); fl = torch.randn(1, 2, 18, 32, device="cuda")
sp = S.Softsplat()
with torch.no_grad():
    for _ in range(50): sp(x, fl)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000): sp(x, fl)
    pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])

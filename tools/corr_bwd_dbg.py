import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import fldr_vfi_b200._lib as L
from oracle import corr_oracle as co, synth
ext = L.ext()
which = sys.argv[1] if len(sys.argv) > 1 else "both"
for (B, Cc, H, W) in [(2, 32, 12, 32), (2, 40, 12, 20), (2, 196, 5, 8), (2, 32, 80, 128)]:
    f1 = synth.features(B, Cc, H, W, seed=3); f2 = synth.features(B, Cc, H, W, seed=5)
    go = synth.grad((B, 81, H, W), seed=4)
    g1, g2 = ext.corr81_bwd(f1.cuda(), f2.cuda(), go.cuda(), which in ("both", "first"), which in ("both", "second"))
    torch.cuda.synchronize()
    e1 = float((g1.cpu() - co.correlation_grad_first(f2, go)).abs().max()) if g1 is not None else -1
    e2 = float((g2.cpu() - co.correlation_grad_second(f1, go)).abs().max()) if g2 is not None else -1
    print("bwd ok", which, B, Cc, H, W, e1, e2, flush=True)

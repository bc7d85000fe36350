import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
from oracle import synth
ext = L.ext()
B, Cc, H, W = 64, 32, 128, 128
f1 = synth.features(B, Cc, H, W, seed=3).cuda(); f2 = synth.features(B, Cc, H, W, seed=5).cuda()
go = synth.grad((B, 81, H, W), seed=4).cuda()
for _ in range(3):
    g1, g2 = ext.corr81_bwd(f1, f2, go, True, True)
torch.cuda.synchronize()

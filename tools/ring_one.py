"""One 4K image splat through the ring kernel (for ncu).  python tools/ring_one.py [F1|F2|F0] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.softSplat as S
from oracle import synth
H, W = 2304, 4096
reg = sys.argv[1] if len(sys.argv) > 1 else "F1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = synth.image(1, 3, H, W, seed=56).cuda(); z = synth.metric(1, H, W, seed=58).cuda()
fl = synth.flow(1, H, W, reg, seed=57).cuda()
for _ in range(n):
    y = S.FunctionSoftsplat(x, fl, z, "softmax")
torch.cuda.synchronize()
print("ok", float(y.mean()))

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.softSplat as S
from oracle import synth
x = synth.image(1, 3, 2304, 4096, seed=71).cuda(); f = synth.flow(1, 2304, 4096, "F1", seed=72).cuda(); z = synth.metric(1, 2304, 4096, seed=73).cuda()
for _ in range(3): S.FunctionSoftsplat(x, f, z, "softmax")
torch.cuda.synchronize()

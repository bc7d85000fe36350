import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
import fldr_vfi_b200.softSplat as S
from oracle import synth
lib = L.lib()
x = synth.image(1, 3, 2304, 4096, seed=71).cuda(); f = synth.flow(1, 2304, 4096, "F1", seed=72).cuda(); z = synth.metric(1, 2304, 4096, seed=73).cuda()
for snake in (0, 1):
    lib.fldr_set_option(b"splat_snake", snake)
    for _ in range(2): S.FunctionSoftsplat(x, f, z, "softmax")
torch.cuda.synchronize()

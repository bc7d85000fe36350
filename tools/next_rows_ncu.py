"""One call of every next-row kernel at the padded 4K frame pair, for ncu: python tools/next_rows_ncu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.blend as Bl
import fldr_vfi_b200.pca as Pc
import fldr_vfi_b200.pyramid as Py
import fldr_vfi_b200.warp as Wp
from oracle import synth
N, C, H, W = 1, 3, 2304, 4096
x0, x1 = synth.image(N, C, H, W, seed=1).cuda(), synth.image(N, C, H, W, seed=2).cuda()
fl = synth.flow(N, H, W, "F1", seed=3).cuda()
g = torch.Generator().manual_seed(5)
mean = (torch.randn(64, generator=g, dtype=torch.float64) * 0.1).cuda()
EV = torch.linalg.qr(torch.randn(64, 64, generator=g, dtype=torch.float64))[0][:16].contiguous().cuda()
mv = (torch.rand(16, generator=g, dtype=torch.float64) + 0.5).cuda()
logits = (synth.grad((N, 6, H, W), seed=77) * 3.0).cuda()
extra = [synth.image(N, C, H, W, seed=78 + k).cuda() for k in range(4)]
tv = torch.full((N, 1, 1, 1), 0.5, device="cuda")
T = torch.ones(1, dtype=torch.float64, device="cuda")
frames = torch.stack([x0, x1], 2).contiguous()
with torch.no_grad():
    for _ in range(2):
        Wp.bwarp(x1, fl, True)
        Wp.splat_metric(x0, x1, fl, -1.894)
        Bl.occ_blend(logits, T, tv, *extra, x0, x1)
        Pc.pca_features(torch.cat([x0[0], x1[0]], 0), mean, EV, mv, out_dtype=torch.float32)
        Py.input_pyramid(frames, [8, 16, 32, 64, 128, 256], 5)
torch.cuda.synchronize()

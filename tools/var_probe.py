"""Image / feature splat timings for the library selected by FLDR_B200_LIB (A/B of kernel variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200.softSplat as S
from oracle import synth
for (N, C, h, w, reg) in [(1, 3, 2304, 4096, "F1"), (1, 3, 2304, 4096, "F2"), (32, 3, 512, 512, "F1"), (1, 48, 288, 512, "F1"), (32, 48, 64, 64, "F1")]:
    x = (synth.image(N, C, h, w, seed=71) if C == 3 else synth.features(N, C, h, w, seed=71)).cuda()
    f = synth.flow(N, h, w, reg, seed=72).cuda()
    z = synth.metric(N, h, w, seed=73).cuda() if C == 3 else None
    fn = lambda: S.FunctionSoftsplat(x, f, z, "softmax")
    fn(); ts = []
    for _ in range(15):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)
        a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 500)
    print(f"({N},{C},{h},{w}) {reg}: {sorted(ts)[7]:.1f} us", flush=True)

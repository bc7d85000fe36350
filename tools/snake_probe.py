"""4K image splat: row order of the three passes (zero fill front to back, scatter bottom-up, normalise top-down) vs the default."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fldr_vfi_b200._lib as L
import fldr_vfi_b200.softSplat as S
from oracle import synth
lib = L.lib()
for (N, C, h, w, reg) in [(1, 3, 2304, 4096, "F1"), (1, 3, 2304, 4096, "F0"), (1, 3, 2304, 4096, "F2"), (1, 3, 1152, 2048, "F1"), (32, 3, 512, 512, "F1"), (1, 48, 288, 512, "F1")]:
    x = (synth.image(N, C, h, w, seed=71) if C == 3 else synth.features(N, C, h, w, seed=71)).cuda()
    f = synth.flow(N, h, w, reg, seed=72).cuda()
    z = synth.metric(N, h, w, seed=73).cuda() if C == 3 else None
    fn = lambda: S.FunctionSoftsplat(x, f, z, "softmax")
    out = []
    ref = None
    for snake, pf in ((0, 0), (1, 0), (1, -1)):
        lib.fldr_set_option(b"splat_snake", snake)
        lib.fldr_set_option(b"splat_pf_rows", pf)
        y = fn()
        if ref is None: ref = y
        ts = []
        for _ in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)
            a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 500)
        out.append(f"snake={snake} pf={pf}: {sorted(ts)[7]:.1f} us (diff {float((y - ref).abs().max()):.1e})")
    lib.fldr_set_option(b"splat_snake", 1); lib.fldr_set_option(b"splat_pf_rows", 0)
    print(f"({N},{C},{h},{w}) {reg}: " + " | ".join(out), flush=True)

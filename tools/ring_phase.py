"""Unbounded-ring run (phases Z, S, N strictly in sequence): raw per-phase throughput of the ring kernel's pipeline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ring_probe as P
from oracle import synth
lib, S = P.lib, P.S
W = 4096
for H in (512, 1024):
    x = synth.image(1, 3, H, W, seed=56).cuda(); z = synth.metric(1, H, W, seed=58).cuda()
    for reg in ("F1", "F0"):
        fl = synth.flow(1, H, W, reg, seed=57).cuda() if reg != "F0" else torch.zeros(1, 2, H, W, device="cuda")
        lib.fldr_set_option(b"splat_stream", 0)
        ref = S.FunctionSoftsplat(x, fl, z, "softmax")
        w_med, _ = P.timeit(lambda: S.FunctionSoftsplat(x, fl, z, "softmax"), iters=10)
        lib.fldr_set_option(b"splat_stream", 1)
        lib.fldr_set_option(b"splat_ring_mb", 68)
        out, flag, info = P.raw_call(x, fl, z)
        nbad, _ = P.close(out, ref)
        med, mn = P.timeit(lambda: S.FunctionSoftsplat(x, fl, z, "softmax"), iters=10)
        alg = 4 * H * W * 9
        print(f"H={H} {reg}: whole-frame {w_med:.1f} us | ring reach={info[1]} rows={info[4]} items={info[5]} flag={flag} bad={nbad}: {med:.1f} us ({mn:.1f} min) frac {alg/med/1e3/6549.1:.3f}", flush=True)

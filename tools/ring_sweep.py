"""Lag / ring-size sweep of the streaming splat kernel at 4K (GPU).  python tools/ring_sweep.py 'mb:lag,mb:lag,...'"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ring_probe as P
from oracle import synth
lib, S = P.lib, P.S
H, W = 2304, 4096
x = synth.image(1, 3, H, W, seed=56).cuda(); z = synth.metric(1, H, W, seed=58).cuda()
flows = {"F1": synth.flow(1, H, W, "F1", seed=57).cuda(), "F0": torch.zeros(1, 2, H, W, device="cuda")}
alg = 4 * H * W * 9
lib.fldr_set_option(b"splat_stream", 0)
refs = {k: S.FunctionSoftsplat(x, fl, z, "softmax") for k, fl in flows.items()}
lib.fldr_set_option(b"splat_stream", 1)
for item in sys.argv[1].split(","):
    mb, lag = (int(v) for v in item.split(":"))
    lib.fldr_set_option(b"splat_ring_mb", mb); lib.fldr_set_option(b"splat_lag", lag)
    for k, fl in flows.items():
        out, flag, info = P.raw_call(x, fl, z)
        nbad, emax = P.close(out, refs[k])
        med, mn = P.timeit(lambda: S.FunctionSoftsplat(x, fl, z, "softmax"), iters=10)
        print(f"  lib={os.path.basename(os.environ.get('FLDR_B200_LIB','default'))} mb={mb} lag={lag} {k}: reach={info[1]} rows={info[4]} flag={flag} bad={nbad} | {med:.1f} us ({mn:.1f} min) frac {alg/med/1e3/6549.1:.3f}", flush=True)

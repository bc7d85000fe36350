"""Per-role wait/busy cycle counters of the ring kernel (library built with -DFLDR_RING_STATS).  python tools/ring_stats.py mb:lag[,..] [F1|F0]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ring_probe as P
from oracle import synth
lib, L = P.lib, P.L
H, W = 2304, 4096
x = synth.image(1, 3, H, W, seed=56).cuda(); z = synth.metric(1, H, W, seed=58).cuda()
reg = sys.argv[2] if len(sys.argv) > 2 else "F1"
fl = synth.flow(1, H, W, reg, seed=57).cuda() if reg != "F0" else torch.zeros(1, 2, H, W, device="cuda")
names = ["cons wait full", "cons S busy", "cons N/Z busy", "gate poll S", "gate poll N", "gate wait posted", "prod wait free", "prod wait stage",
         "prod wait ticket", "sig fence+signal", "sig wait done", "S poll spins", "N poll spins"]
for item in sys.argv[1].split(","):
    mb, lag = (int(v) for v in item.split(":"))
    lib.fldr_set_option(b"splat_ring_mb", mb); lib.fldr_set_option(b"splat_lag", lag)
    info = P.plan(3, 1, 3, H, W, True)
    ws = torch.zeros(info[6], dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x)
    for it in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st = lib.fldr_splat_fwd(3, L.ptr(x), L.strides(x), L.ptr(fl), L.strides(fl), L.ptr(z), L.strides(z), L.ptr(out), None, 1, 3, H, W,
                                L.ptr(ws), info[6], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3
    base = info[3] - 32 * 4 + 40 * 4
    st = ws[base:base + 13 * 8].view(torch.int64).cpu().tolist()
    ctas = 148 * 4
    print(f"mb={mb} lag={lag} {reg}: {us:.1f} us; per-role cycles / (CTAs*warps) in us @1.9GHz:")
    for i, n in enumerate(names):
        if i >= 11: print(f"    {n:18s} {st[i]} total ({st[i]/9216:.1f} per item)")
        else:
            warps = 4 if i < 3 else 1
            print(f"    {n:18s} {st[i]/1.9e3/(ctas*warps):8.1f} us per warp")

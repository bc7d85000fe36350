"""Per-item event timeline of the first CTAs of the ring kernel (library built with -DFLDR_RING_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ring_probe as P
from oracle import synth
lib, L = P.lib, P.L
H, W = (int(sys.argv[1]) if len(sys.argv) > 1 else 2304), 4096
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib.fldr_set_option(b"splat_ring_mb", mb)
x = synth.image(1, 3, H, W, seed=56).cuda(); z = synth.metric(1, H, W, seed=58).cuda()
fl = synth.flow(1, H, W, "F1", seed=57).cuda()
info = P.plan(3, 1, 3, H, W, True)
ws = torch.zeros(info[6], dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
for it in range(2):
    st = lib.fldr_splat_fwd(3, L.ptr(x), L.strides(x), L.ptr(fl), L.strides(fl), L.ptr(z), L.strides(z), L.ptr(out), None, 1, 3, H, W,
                            L.ptr(ws), info[6], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
NS = (H + 7) // 8
nZs = min(info[4] // 8, NS)
words = 64 + NS + nZs + NS
words += words & 1
base = info[3] - 32 * 4 + words * 4
tr = ws[base:base + 8 * 128 * 8 * 8].view(torch.int64).cpu().view(8, 128, 8)
names = {0: "exit", 1: "Z", 2: "S", 3: "N"}
for blk in (0, 5):
    t = tr[blk]
    t0 = int(t[0, 0])
    print(f"CTA {blk}: item kind | claim_start claim_end posted | cons_start cons_end | signalled   (us since first claim, 1.9 GHz)")
    for k in range(64):
        if int(t[k, 0]) == 0: break
        us = lambda v: (int(v) - t0) / 1900.0 if int(v) else -1
        print(f"  {k:3d} {names.get(int(t[k,6]),'?'):4s} | {us(t[k,0]):7.1f} {us(t[k,1]):7.1f} {us(t[k,2]):7.1f} | {us(t[k,3]):7.1f} {us(t[k,4]):7.1f} | {us(t[k,5]):7.1f}")

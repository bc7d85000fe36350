"""Time PWC-Net's cost-volume step at the native-4K C = 32 / 64 levels: correlation + leaky_relu(0.1) + placement in the
decoder's concatenation buffer (PWCNet.py:146-160), fused (FunctionCorrelationLeakyReLU(out=buf)) vs the drop-in
correlation followed by the torch operators the reference runs.   python tools/corr_act_probe.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fldr_vfi_b200.correlation as C   # noqa: E402
from oracle import synth                # noqa: E402  (input generation only)


def timeit(fn, reps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    for name, B, Cc, H, W in (("C32", 2, 32, 576, 1024), ("C64", 2, 64, 288, 512)):
        f1 = synth.features(B, Cc, H, W, seed=1).cuda()
        f2 = synth.features(B, Cc, H, W, seed=2).cuda()
        flow = torch.zeros(B, 2, H, W, device="cuda")
        feat = torch.zeros(B, 2, H, W, device="cuda")
        buf = torch.empty(B, 81 + Cc + 2 + 2, H, W, device="cuda")
        with torch.no_grad():
            def fused():
                C.FunctionCorrelationLeakyReLU(tensorFirst=f1, tensorSecond=f2, out=buf)
                buf[:, 81:81 + Cc] = f1; buf[:, 81 + Cc:83 + Cc] = flow; buf[:, 83 + Cc:] = feat

            def unfused():
                vol = torch.nn.functional.leaky_relu(C.FunctionCorrelation(tensorFirst=f1, tensorSecond=f2), negative_slope=0.1)
                return torch.cat([vol, f1, flow, feat], 1)
            t_f, t_u = timeit(fused), timeit(unfused)
            t_c = timeit(lambda: C.FunctionCorrelation(tensorFirst=f1, tensorSecond=f2))
        print(json.dumps({"level": name, "shape": f"{B}x{Cc}x{H}x{W}", "corr_only_us": round(t_c, 1), "fused_leaky_into_concat_us": round(t_f, 1),
                          "corr_then_torch_leaky_and_cat_us": round(t_u, 1), "saved_us": round(t_u - t_f, 1)}), flush=True)


if __name__ == "__main__":
    main()

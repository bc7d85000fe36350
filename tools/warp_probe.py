"""Time the backward warp / splat metric row at the 4K image shape (2304x4096, C=3) and at the flow-warp shape (C=2),
next to the reference's own method (fLDRnet.py:546-581, lifted from baseline/_ref by baseline/ref_src.py) on the same GPU.

    python tools/warp_probe.py          # prints one JSON line per measurement
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fldr_vfi_b200.warp as Wp   # noqa: E402
from baseline import ref_src       # noqa: E402  (the reference's own bwarp text, from baseline/_ref)
from oracle import synth           # noqa: E402  (input generation only)

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


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
    H, W = 2304, 4096
    x0 = synth.image(1, 3, H, W, seed=0).cuda()
    x1 = synth.image(1, 3, H, W, seed=1).cuda()
    fl = synth.flow(1, H, W, "F1", seed=2).cuda()
    fl2 = synth.flow(1, H, W, "F1", seed=3).cuda()
    alpha = -1.894
    px = H * W
    # the reference's method with its grid / ones tensors allocated on the device (as written it builds them on the CPU
    # and copies them over per call - that version is timed too, as "as_written")
    reference_bwarp = ref_src.bwarp(x0.device, create_on_device=True)
    as_written = ref_src.bwarp(x0.device)
    with torch.no_grad():
        rows = [
            ("bwarp C=3 (image)", 4 * px * (3 + 2 + 3), lambda: Wp.bwarp(x1, fl, True), lambda: reference_bwarp(x1, fl, True)),
            ("bwarp C=2 (flow by flow, fLDRnet.py:474)", 4 * px * (2 + 2 + 2), lambda: Wp.bwarp(fl2, fl, True), lambda: reference_bwarp(fl2, fl, True)),
            ("splat_metric C=3 (fLDRnet.py:442-443)", 4 * px * (3 + 3 + 2 + 1), lambda: Wp.splat_metric(x0, x1, fl, alpha),
             lambda: torch.mean(alpha * torch.abs(x0 - reference_bwarp(x1, fl, True)), dim=1, keepdim=True)),
        ]
        for what, nbytes, ours, ref in rows:
            t_o, t_r = timeit(ours), timeit(ref, reps=10)
            print(json.dumps({"op": what, "shape": f"1x{H}x{W}", "algorithmic_MB": round(nbytes / 1e6, 1), "ours_us": round(t_o, 1),
                              "ours_GBps": round(nbytes / t_o / 1e3, 1), "frac_of_hbm_peak": round(nbytes / t_o / 1e3 / PEAK, 3),
                              "torch_reference_ops_us": round(t_r, 1), "speedup": round(t_r / t_o, 1)}), flush=True)
        t_w = timeit(lambda: as_written(x1, fl, True), reps=5)
        print(json.dumps({"op": "bwarp C=3 (image), reference method as written (CPU-built grid and ones)", "us": round(t_w, 1)}), flush=True)


if __name__ == "__main__":
    main()

"""Time the occlusion softmax + image synthesis row at the 4K shape next to the reference's own statements
(fLDRnet.py:510-524, lifted from baseline/_ref by baseline/ref_src.py) on the same GPU.    python tools/blend_probe.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fldr_vfi_b200.blend as Bl   # noqa: E402
from baseline import ref_src        # noqa: E402  (the reference's own statements, from baseline/_ref)
from oracle import synth            # noqa: E402  (input generation only)

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
    H, W, C = 2304, 4096, 3
    imgs = [synth.image(1, C, H, W, seed=500 + k).cuda() for k in range(6)]
    refine = (synth.grad((1, 6, H, W), seed=510) * 3.0).cuda()
    tv = torch.full((1, 1, 1, 1), 0.5, device="cuda")
    T = torch.ones(1, dtype=torch.float64, device="cuda")
    nbytes = H * W * (6 * 4 + 6 * C * 4 + C * 8)
    reference_ops = ref_src.blend()
    x_l = torch.stack([imgs[4], imgs[5]], 2)
    with torch.no_grad():
        t_o = timeit(lambda: Bl.occ_blend(refine, T, tv, *imgs))
        t_r = timeit(lambda: reference_ops(refine, T, tv, imgs[0], imgs[1], imgs[2], imgs[3], x_l), reps=10)
    print(json.dumps({"op": "occ_blend C=3 (fLDRnet.py:510-524), float64", "shape": f"1x{H}x{W}", "algorithmic_MB": round(nbytes / 1e6, 1),
                      "ours_us": round(t_o, 1), "ours_GBps": round(nbytes / t_o / 1e3, 1), "frac_of_hbm_peak": round(nbytes / t_o / 1e3 / PEAK, 3),
                      "torch_reference_ops_us": round(t_r, 1), "speedup": round(t_r / t_o, 1)}), flush=True)


if __name__ == "__main__":
    main()

"""Timing of the block-PCA feature extraction at the padded 4K frame pair (tools; CUDA events, device time)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fldr_vfi_b200.pca as P

g = torch.Generator().manual_seed(5)
mean = (torch.randn(64, generator=g, dtype=torch.float64) * 0.1).cuda()
EV = torch.linalg.qr(torch.randn(64, 64, generator=g, dtype=torch.float64))[0][:16].contiguous().cuda()
mv = (torch.rand(16, generator=g, dtype=torch.float64) + 0.5).cuda()
for (chan, H, W) in [(6, 2304, 4096), (6, 1152, 2048), (192, 512, 512)]:
    im = torch.rand(chan, H, W, device="cuda") * 2 - 1
    for dt in (torch.float32, torch.float64):
        fn = lambda: P.pca_features(im, mean, EV, mv, out_dtype=dt)
        with torch.no_grad():
            fn()
            ts = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(400000)
                a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 2)
        ms = sorted(ts)[len(ts) // 2]
        fl = 2.0 * chan * H * W * 16
        print(f"({chan},{H},{W}) out {str(dt)[6:]}: {ms*1e3:.1f} us  {fl/ms/1e9:.2f} TFLOP/s f64")

"""Seeded synthetic inputs for parity tests and the bench.

Everything is generated on the CPU in fp32 from ``torch.Generator().manual_seed(s)`` so the
oracle, the reference kernels and the CUDA path see identical bits.  (Input generation only:
no reference arithmetic lives here, so ``bench.py`` may import it for its inputs.)
"""
import torch
import torch.nn.functional as F


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(int(seed))


def image(N, C, H, W, seed=0):
    """Band-limited noise in [-1, 1]: randn at 1/16 res, bicubic up, tanh."""
    h, w = max(H // 16, 2), max(W // 16, 2)
    x = torch.randn(N, C, h, w, generator=_gen(seed))
    return torch.tanh(F.interpolate(x, size=(H, W), mode="bicubic", align_corners=False)).contiguous()


def flow(N, H, W, regime="F1", seed=1):
    """Flow regimes: F0 zero, F1 smooth (sigma 16 px at W=4096, scaled with W), F2 iid U(-64,64)
    (scaled with W/4096, at least +-4), F3 converge to centre +-2, FB border (25 % of targets out)."""
    g = _gen(seed)
    if regime == "F0":
        return torch.zeros(N, 2, H, W)
    if regime == "F1":
        sigma = 16.0 * W / 4096.0
        h, w = max(H // 64, 2), max(W // 64, 2)
        lo = torch.randn(N, 2, h, w, generator=g) * sigma
        return F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False).contiguous()
    if regime == "F2":
        a = max(64.0 * W / 4096.0, 4.0)
        return (torch.rand(N, 2, H, W, generator=g) * 2 - 1) * a
    if regime == "F3":
        gx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(N, 1, H, W)
        gy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(N, 1, H, W)
        tgt = torch.cat([(W / 2 - gx), (H / 2 - gy)], 1)
        return (tgt + (torch.rand(N, 2, H, W, generator=g) * 4 - 2)).contiguous()
    if regime == "FB":
        # uniform shift by a quarter frame plus jitter: ~25 % of targets leave the frame
        base = torch.tensor([W / 4.0, 0.0]).view(1, 2, 1, 1)
        return (base + torch.randn(N, 2, H, W, generator=g) * 1.5).contiguous()
    raise ValueError(regime)


def metric(N, H, W, seed=2):
    """z = z_alpha * mean|diff| with the shipped checkpoint's z_alpha ~ -1.894 -> z in [-3.79, 0]."""
    return (-1.894 * 2.0 * torch.rand(N, 1, H, W, generator=_gen(seed))).contiguous()


def features(N, C, H, W, seed=3):
    return F.leaky_relu(torch.randn(N, C, H, W, generator=_gen(seed)), 0.1).contiguous()


def grad(shape, seed=4):
    return torch.randn(*shape, generator=_gen(seed)).contiguous()

"""GPU parity tests of the device-side input pyramid (main.py:855-856) through the C-ABI: golden vectors from
the reference's own expression evaluated on the CPU, the oracle on seeded shapes incl. the padded 4K frame pair, properties."""
import numpy as np
import pytest
import torch

from oracle import pyramid_oracle as po
from oracle import synth
from util import load_golden

pytestmark = pytest.mark.gpu
TOL = 2e-6          # float32, 16 taps, x max|frame|: fma / summation-order differences only (tap positions and weights are exact)


@pytest.mark.parametrize("name", ["pyramid_5lv", "pyramid_noise_b2", "pyramid_ac", "pyramid_odd"])
def test_pyramid_vs_golden(cuda_lib, name):
    import fldr_vfi_b200.pyramid as P
    g = load_golden(name)
    frames = g["frames"].cuda()
    n = int(g["n_levels"])
    levels = P.input_pyramid(frames, [int(s) for s in g["scales"]], n, align_corners=bool(int(g["align_corners"])))
    assert len(levels) == n + 1 and levels[0] is frames
    mag = max(1.0, float(g["frames"].abs().max()))
    for i in range(1, n + 1):
        ref = g[f"level{i}"]
        assert tuple(levels[i].shape) == tuple(ref.shape) and levels[i].dtype == torch.float32
        assert float((levels[i].cpu() - ref).abs().max()) <= TOL * mag, (name, i)


@pytest.mark.parametrize("B,T,H,W,n", [(1, 2, 96, 384, 5), (2, 2, 64, 160, 5), (1, 1, 40, 72, 3), (1, 2, 8, 4, 2), (1, 2, 32, 132, 2)])
def test_pow2_path_vs_oracle_seeded(cuda_lib, B, T, H, W, n):
    """Single-pass kernel on ragged blocks: W not a multiple of 128, H not a multiple of 32, frames smaller than one block."""
    import fldr_vfi_b200.pyramid as P
    frames = torch.randn(B, 3, T, H, W, generator=torch.Generator().manual_seed(11))
    scales = [8 << i for i in range(n + 1)]
    ref = po.input_pyramid(frames.numpy(), scales, n)
    got = P.input_pyramid(frames.cuda(), scales, n)
    mag = float(frames.abs().max())
    for i in range(1, n + 1):
        assert tuple(got[i].shape) == ref[i].shape
        assert float(np.abs(got[i].cpu().numpy() - ref[i]).max()) <= TOL * mag, (i, got[i].shape)


@pytest.mark.parametrize("factors,ac", [([0.5, 0.25], True), ([1 / 3, 0.2], False), ([0.25, 0.125], False), ([0.75], False), ([2.0], False)])
def test_generic_path_vs_oracle(cuda_lib, factors, ac):
    """align_corners, factors that are not 1/2^k, level lists that skip 1/2 (no single-pass form), and an upscale."""
    import fldr_vfi_b200.pyramid as P
    x = torch.randn(5, 37, 53, generator=torch.Generator().manual_seed(12))
    got = P.bicubic_levels(x.cuda(), factors, align_corners=ac)
    for f, o in zip(factors, got):
        ref = po.bicubic_resize(x.numpy(), f, ac)
        assert tuple(o.shape) == ref.shape
        assert float(np.abs(o.cpu().numpy() - ref).max()) <= TOL * float(x.abs().max())


def test_generic_and_pow2_paths_agree_and_views(cuda_lib):
    """A row-strided view (crop of a wider buffer) goes through the same kernels; an unaligned crop falls to the generic path
    and must agree with the single-pass result."""
    import fldr_vfi_b200.pyramid as P
    big = torch.randn(6, 64, 264, generator=torch.Generator().manual_seed(13)).cuda()
    aligned, shifted = big[:, :, 4:260], big[:, :, 1:257]
    for v in (aligned, shifted):
        got = P.bicubic_levels(v, [0.5, 0.25, 0.125])
        ref = [po.bicubic_resize(v.cpu().numpy(), f) for f in (0.5, 0.25, 0.125)]
        for o, r in zip(got, ref):
            assert float(np.abs(o.cpu().numpy() - r).max()) <= TOL * float(big.abs().max())


def test_pyramid_4k_frame_pair(cuda_lib):
    """BASELINE configs[2] shape: two RGB frames padded to 2304 x 4096, five levels (--test5scales).  Level 1..5 against the
    oracle (seconds on the CPU) and two size-independent properties: a constant frame stays constant, and the pyramid is linear."""
    import fldr_vfi_b200.pyramid as P
    frames = synth.image(1, 6, 2304, 4096, seed=3).reshape(1, 3, 2, 2304, 4096)
    scales = [8, 16, 32, 64, 128, 256]
    got = P.input_pyramid(frames.cuda(), scales, 5)
    ref = po.input_pyramid(frames.numpy(), scales, 5)
    for i in range(1, 6):
        assert tuple(got[i].shape) == (1, 3, 2, 2304 >> i, 4096 >> i)
        assert float(np.abs(got[i].cpu().numpy() - ref[i]).max()) <= TOL
    const = torch.full((1, 3, 2, 2304, 4096), 0.625, device="cuda")
    for lv in P.input_pyramid(const, scales, 5)[1:]:
        assert float((lv - 0.625).abs().max()) <= 1e-7
    other = torch.randn(1, 3, 2, 2304, 4096, device="cuda")
    a, b, ab = P.input_pyramid(frames.cuda(), scales, 5), P.input_pyramid(other, scales, 5), P.input_pyramid(frames.cuda() + 2 * other, scales, 5)
    for i in range(1, 6):
        assert float((ab[i] - (a[i] + 2 * b[i])).abs().max()) <= 2e-5


def test_pyramid_errors_and_empty(cuda_lib):
    import fldr_vfi_b200.pyramid as P
    x = torch.zeros(1, 3, 2, 16, 16)
    with pytest.raises(NotImplementedError):
        P.input_pyramid(x, [8, 16], 1)                                   # CPU tensor: no fallback
    with pytest.raises(TypeError):
        P.input_pyramid(x.cuda().double(), [8, 16], 1)
    with pytest.raises(RuntimeError):
        P.bicubic_levels(x.cuda(), [1 / 64])                             # a level with zero pixels
    assert P.input_pyramid(x.cuda(), [8], 0)[0].shape == x.shape         # no levels: the frames themselves
    empty = P.bicubic_levels(torch.zeros(0, 16, 16, device="cuda"), [0.5])
    assert tuple(empty[0].shape) == (0, 8, 8)

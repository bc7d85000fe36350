"""Seeded random-shape parity sweep on the GPU: every mode, with / without metric, ragged sizes (W % 4 != 0 takes the
plain-load correlation kernel, W % 4 == 0 the TMA one), channel counts that do not fill a quad or a channel chunk,
strided input views, batch > 1 - forward and all gradients against the CPU oracle."""
import random

import pytest
import torch

from oracle import corr_oracle as co
from oracle import splat_oracle as so
from oracle import synth
from util import assert_corr_close, assert_splat_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_splat_random_case(cuda_lib, seed):
    import fldr_vfi_b200.softSplat as S
    rng = random.Random(1000 + seed)
    N, C = rng.randint(1, 3), rng.randint(1, 9)
    H, W = rng.randint(1, 40), rng.randint(1, 70)
    mode = rng.choice(["summation", "average", "linear", "softmax", "softmax"])
    with_metric = mode == "linear" or (mode == "softmax" and rng.random() < 0.6)
    regime = rng.choice(["F0", "F1", "F2", "F3", "FB"])
    scale = rng.choice([0.3, 1.0, 4.0]) if regime in ("F1", "F2") else 1.0
    x_full = synth.features(N, C + 2, H, W, seed=seed * 7 + 1)
    x = x_full[:, 1:C + 1]                                   # channel-sliced view (non-contiguous for N > 1)
    fl_full = synth.flow(N, H, W, regime, seed=seed * 7 + 2) * scale
    fl4 = torch.cat([fl_full, fl_full.flip(1)], 1)
    fl = fl4[:, :2]                                          # view into a 4-channel flow (fLDRnet.py:386)
    z = synth.metric(N, H, W, seed=seed * 7 + 3) if with_metric else None
    g = synth.grad((N, C, H, W), seed=seed * 7 + 4)
    xd = x_full.cuda()[:, 1:C + 1].requires_grad_(True)
    fd = fl4.cuda()[:, :2].requires_grad_(True)
    zd = None if z is None else z.cuda().requires_grad_(True)
    y = S.FunctionSoftsplat(xd, fd, zd, mode)
    wrt = [xd, fd] + ([zd] if zd is not None and mode in ("linear", "softmax") else [])
    grads = torch.autograd.grad(y, wrt, g.cuda())
    yo, gi, gf, gz = so.function_softsplat_grads(x, fl, z if mode in ("linear", "softmax") else None, mode, g)
    _, gi64, gf64, gz64 = so.function_softsplat_grads(x.double(), fl.double(), None if z is None or mode not in ("linear", "softmax") else z.double(), mode, g.double())
    what = f"seed {seed}: {mode} N{N} C{C} {H}x{W} {regime} metric={with_metric}"
    assert_splat_close(y, yo, what + " out", mag=None if mode in ("summation", "linear") else 1.0)
    assert_splat_close(grads[0], gi, what + " grad_input", cond=gi - gi64)
    assert_splat_close(grads[1], gf, what + " grad_flow", cond=gf - gf64)
    if len(grads) == 3:
        assert_splat_close(grads[2], gz, what + " grad_metric", cond=gz - gz64)


@pytest.mark.parametrize("seed", range(16))
def test_correlation_random_case(cuda_lib, seed):
    import fldr_vfi_b200.correlation as Cm
    rng = random.Random(2000 + seed)
    B, C = rng.randint(1, 3), rng.choice([1, 3, 7, 8, 9, 16, 31, 33, 40, 64, 100])
    H = rng.randint(1, 30)
    W = rng.choice([rng.randint(1, 50), 4 * rng.randint(1, 16)])        # ragged or multiple of 4
    f1 = synth.features(B, C, H, W, seed=seed * 5 + 1)
    f2 = synth.features(B, C, H, W, seed=seed * 5 + 2)
    g = synth.grad((B, 81, H, W), seed=seed * 5 + 3)
    a, b = f1.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
    out = Cm.FunctionCorrelation(tensorFirst=a, tensorSecond=b)
    g1, g2 = torch.autograd.grad(out, [a, b], g.cuda())
    what = f"seed {seed}: B{B} C{C} {H}x{W}"
    assert_corr_close(out, co.correlation_fwd(f1, f2), co.correlation_fwd(f1.abs(), f2.abs()), what + " fwd")
    assert_corr_close(g1, co.correlation_grad_first(f2, g), co.correlation_grad_first(f2.abs(), g.abs()), what + " gradFirst")
    assert_corr_close(g2, co.correlation_grad_second(f1, g), co.correlation_grad_second(f1.abs(), g.abs()), what + " gradSecond")

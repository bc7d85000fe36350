"""GPU parity tests for the 81-channel correlation: CUDA path (through the C-ABI) vs the CPU oracle, the golden
vectors of the reference kernel text, PWC pyramid shapes, and properties at the native 4K size."""
import pytest
import torch

from oracle import corr_oracle as co
from oracle import synth
from util import assert_corr_close, load_golden

pytestmark = pytest.mark.gpu

CORR_CASES = ["corr_c40", "corr_c32_odd", "corr_c196_tiny"]


def _mod(cuda_lib):
    import fldr_vfi_b200.correlation as C
    return C


def _check(Cm, f1, f2, gout, what):
    f1d = f1.cuda().requires_grad_(True)
    f2d = f2.cuda().requires_grad_(True)
    out = Cm.FunctionCorrelation(tensorFirst=f1d, tensorSecond=f2d)
    assert out.shape == (f1.shape[0], 81, f1.shape[2], f1.shape[3]) and out.is_contiguous()
    assert_corr_close(out, co.correlation_fwd(f1, f2), co.correlation_fwd(f1.abs(), f2.abs()), what + " fwd")
    g1, g2 = torch.autograd.grad(out, [f1d, f2d], gout.cuda())
    assert_corr_close(g1, co.correlation_grad_first(f2, gout), co.correlation_grad_first(f2.abs(), gout.abs()), what + " gradFirst")
    assert_corr_close(g2, co.correlation_grad_second(f1, gout), co.correlation_grad_second(f1.abs(), gout.abs()), what + " gradSecond")


@pytest.mark.parametrize("name", CORR_CASES)
def test_vs_golden(cuda_lib, name):
    Cm = _mod(cuda_lib)
    g = load_golden(name)
    f1, f2 = g["first"], g["second"]
    f1d = f1.cuda().requires_grad_(True)
    f2d = f2.cuda().requires_grad_(True)
    out = Cm.ModuleCorrelation()(f1d, f2d)
    assert_corr_close(out, g["out"], co.correlation_fwd(f1.abs(), f2.abs()), name + " fwd")
    g1, g2 = torch.autograd.grad(out, [f1d, f2d], g["grad_out"].cuda())
    assert_corr_close(g1, g["grad_first"], co.correlation_grad_first(f2.abs(), g["grad_out"].abs()), name + " gradFirst")
    assert_corr_close(g2, g["grad_second"], co.correlation_grad_second(f1.abs(), g["grad_out"].abs()), name + " gradSecond")


# cfg2 literal pyramid (4K / 8 -> 320x512, B=2) and ragged shapes
@pytest.mark.parametrize("B,C,H,W", [
    (2, 196, 5, 8), (2, 128, 10, 16), (2, 96, 20, 32), (2, 64, 40, 64), (2, 32, 80, 128),
    (1, 1, 1, 1), (1, 3, 9, 7), (2, 17, 13, 37), (1, 33, 8, 70),
])
def test_vs_oracle_seeded(cuda_lib, B, C, H, W):
    Cm = _mod(cuda_lib)
    f1 = synth.features(B, C, H, W, seed=3)
    f2 = synth.features(B, C, H, W, seed=5)
    gout = synth.grad((B, 81, H, W), seed=4)
    _check(Cm, f1, f2, gout, f"B{B} C{C} {H}x{W}")


def test_needs_input_grad(cuda_lib):
    Cm = _mod(cuda_lib)
    f1 = synth.features(1, 16, 12, 12, seed=1).cuda().requires_grad_(True)
    f2 = synth.features(1, 16, 12, 12, seed=2).cuda()
    out = Cm.FunctionCorrelation(f1, f2)
    out.sum().backward()
    assert f1.grad is not None and f2.grad is None


def test_error_behaviour(cuda_lib):
    Cm = _mod(cuda_lib)
    f = torch.zeros(1, 4, 8, 8)
    with pytest.raises(NotImplementedError):
        Cm.FunctionCorrelation(f, f)                                       # correlation.py:343-344
    with pytest.raises(AssertionError):
        Cm.FunctionCorrelation(f.cuda().permute(0, 1, 3, 2), f.cuda())    # correlation.py:302
    with pytest.raises(TypeError):
        Cm.FunctionCorrelation(f.cuda().double(), f.cuda().double())


def test_native_4k_level_properties_and_crop(cuda_lib):
    """cfg2-native level 2 (B=2, C=32, 544x1024): f1==f2 -> centre channel = mean f1^2; swap symmetry; the oracle at full size (~3 s)."""
    Cm = _mod(cuda_lib)
    B, C, H, W = 2, 32, 544, 1024
    f1 = synth.features(B, C, H, W, seed=61)
    f2 = synth.features(B, C, H, W, seed=62)
    f1d, f2d = f1.cuda(), f2.cuda()
    same = Cm.FunctionCorrelation(f1d, f1d)
    assert float((same[:, 40] - (f1d * f1d).mean(1)).abs().max()) <= 1e-5
    a = Cm.FunctionCorrelation(f1d, f2d)
    bsw = Cm.FunctionCorrelation(f2d, f1d)
    # corr(f1,f2)[dy,dx](y,x) == corr(f2,f1)[-dy,-dx](y+dy,x+dx)
    for (dy, dx) in ((-4, -4), (0, 3), (2, -1), (4, 4)):
        ch, chm = (dy + 4) * 9 + dx + 4, (-dy + 4) * 9 - dx + 4
        ya, xa = slice(max(0, -dy), H - max(0, dy)), slice(max(0, -dx), W - max(0, dx))
        yb, xb = slice(max(0, dy), H - max(0, -dy)), slice(max(0, dx), W - max(0, -dx))
        assert float((a[:, ch, ya, xa] - bsw[:, chm, yb, xb]).abs().max()) <= 2e-6
    assert_corr_close(a, co.correlation_fwd(f1, f2), co.correlation_fwd(f1.abs(), f2.abs()), "native level 2 vs oracle")


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 272, 512), (2, 96, 136, 256), (2, 128, 68, 128), (2, 196, 34, 64)])
def test_native_4k_pyramid_levels_vs_oracle(cuda_lib, B, C, H, W):
    """cfg2-native levels 3..6 at their literal sizes (level 2, C = 32, is the test above): the tile scheduling of C = 64 / 96 and the
    channel-split path of C = 128 / 196, forward against the oracle at full size."""
    Cm = _mod(cuda_lib)
    f1 = synth.features(B, C, H, W, seed=63)
    f2 = synth.features(B, C, H, W, seed=64)
    with torch.no_grad():
        out = Cm.FunctionCorrelation(tensorFirst=f1.cuda(), tensorSecond=f2.cuda())
    assert_corr_close(out, co.correlation_fwd(f1, f2), co.correlation_fwd(f1.abs(), f2.abs()), f"native level C={C}")


@pytest.mark.parametrize("B,C,H,W", [(64, 32, 128, 128), (64, 64, 64, 64), (64, 96, 32, 32), (64, 196, 8, 8)])
def test_cfg5_training_shapes_fwd_bwd(cuda_lib, B, C, H, W):
    """BASELINE configs[4] at the literal batch (B = 64 = 32 crops x 2 directions): the GPU runs the whole batch, forward and both
    gradients; the oracle checks samples 0, 31 and 63 (samples are independent, correlation.py:365,385 launches them one by one)."""
    Cm = _mod(cuda_lib)
    f1 = synth.features(B, C, H, W, seed=65)
    f2 = synth.features(B, C, H, W, seed=66)
    gout = synth.grad((B, 81, H, W), seed=67)
    f1d, f2d = f1.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
    out = Cm.FunctionCorrelation(tensorFirst=f1d, tensorSecond=f2d)
    g1, g2 = torch.autograd.grad(out, [f1d, f2d], gout.cuda())
    for b in (0, 31, 63):
        sl = slice(b, b + 1)
        a, c, g = f1[sl], f2[sl], gout[sl]
        assert_corr_close(out[sl], co.correlation_fwd(a, c), co.correlation_fwd(a.abs(), c.abs()), f"cfg5 C={C} sample {b} fwd")
        assert_corr_close(g1[sl], co.correlation_grad_first(c, g), co.correlation_grad_first(c.abs(), g.abs()), f"cfg5 C={C} sample {b} gradFirst")
        assert_corr_close(g2[sl], co.correlation_grad_second(a, g), co.correlation_grad_second(a.abs(), g.abs()), f"cfg5 C={C} sample {b} gradSecond")


@pytest.mark.parametrize("rows", [0, 2])
def test_backward_kernels_agree(cuda_lib, rows):
    """Both backward kernels (three-rows-per-thread TMA ring, 4-row tile) against the oracle on a ragged multi-chunk shape."""
    Cm = _mod(cuda_lib)
    old = cuda_lib.fldr_get_option(b"corr_bwd_rows")
    cuda_lib.fldr_set_option(b"corr_bwd_rows", rows)
    try:
        for (B, C, H, W) in [(2, 70, 13, 36), (1, 32, 7, 8), (3, 33, 19, 44)]:
            _check(Cm, synth.features(B, C, H, W, seed=21), synth.features(B, C, H, W, seed=22), synth.grad((B, 81, H, W), seed=23),
                   f"bwd rows={rows} B{B} C{C} {H}x{W}")
    finally:
        cuda_lib.fldr_set_option(b"corr_bwd_rows", old)


# ---------------------------------------------------------------- next row: fused leaky-relu + concat placement
@pytest.mark.parametrize("B,C,H,W", [(2, 32, 24, 40), (1, 40, 9, 11), (2, 196, 5, 8), (2, 64, 72, 128), (1, 16, 17, 24)])
def test_leaky_relu_epilogue_and_concat_buffer(cuda_lib, B, C, H, W):
    """leaky_relu(corr, 0.1) (PWCNet.py:146-158) from the fused epilogue equals the activation applied to the oracle's
    volume, dense and when written straight into channels 0..80 of a wider concatenation buffer (PWCNet.py:160)."""
    Cm = _mod(cuda_lib)
    f1 = synth.features(B, C, H, W, seed=11)
    f2 = synth.features(B, C, H, W, seed=12)
    want = torch.nn.functional.leaky_relu(co.correlation_fwd(f1, f2), 0.1)
    scale = co.correlation_fwd(f1.abs(), f2.abs())
    with torch.no_grad():
        got = Cm.FunctionCorrelationLeakyReLU(tensorFirst=f1.cuda(), tensorSecond=f2.cuda())
        assert_corr_close(got, want, scale, "fused leaky dense")
        plain = Cm.FunctionCorrelation(tensorFirst=f1.cuda(), tensorSecond=f2.cuda())
        # (not bit for bit: small levels split the channel range over CTAs and reduce with unordered REDs)
        assert_corr_close(got, torch.nn.functional.leaky_relu(plain, 0.1).cpu(), scale, "fused vs activation of the plain result")
        for extra in (C + 2, 7):                                    # 7: sample stride not a multiple of 4 -> generic kernel
            buf = torch.full((B, 81 + extra, H, W), 123.0, device="cuda")
            view = Cm.FunctionCorrelationLeakyReLU(tensorFirst=f1.cuda(), tensorSecond=f2.cuda(), out=buf)
            assert view.data_ptr() == buf.data_ptr() and view.shape == (B, 81, H, W)
            assert_corr_close(buf[:, :81], want, scale, "fused leaky into concat buffer")
            assert bool((buf[:, 81:] == 123.0).all()), "channels beyond the volume must be left alone"
        same = Cm.FunctionCorrelationLeakyReLU(tensorFirst=f1.cuda(), tensorSecond=f2.cuda(), negative_slope=1.0)
        assert_corr_close(same, plain.cpu(), scale, "slope 1 = plain correlation")


def test_leaky_relu_entry_point_gradients(cuda_lib):
    """PWCNet.py:146-158 under autograd: leaky_relu(correlation) - fused forward, backward through the slope mask and the
    correlation backward kernels - against autograd through the oracle."""
    import torch.nn.functional as F
    Cm = _mod(cuda_lib)
    B, C, H, W = 2, 24, 12, 20
    f1, f2 = synth.features(B, C, H, W, seed=91), synth.features(B, C, H, W, seed=92)
    go = synth.grad((B, 81, H, W), seed=93)
    a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    ref = F.leaky_relu(co.correlation_fwd(a, b), 0.1)
    g1, g2 = torch.autograd.grad(ref, [a, b], go)
    ad, bd = f1.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
    out = Cm.FunctionCorrelationLeakyReLU(tensorFirst=ad, tensorSecond=bd, negative_slope=0.1)
    d1, d2 = torch.autograd.grad(out, [ad, bd], go.cuda())
    assert float((d1.cpu() - g1).abs().max()) < 1e-5 and float((d2.cpu() - g2).abs().max()) < 1e-5
    with pytest.raises(NotImplementedError):             # writing into a caller's concat buffer stays forward-only
        Cm.FunctionCorrelationLeakyReLU(tensorFirst=ad, tensorSecond=bd, out=torch.empty(B, 90, H, W, device="cuda"))

"""GPU parity tests for the splat: CUDA path (through the C-ABI) vs the CPU oracle, the golden vectors of the
reference kernel text, and size-independent properties at the full 4K size."""
import pytest
import torch

from oracle import splat_oracle as so
from oracle import synth
from util import assert_splat_close, load_golden

pytestmark = pytest.mark.gpu

SPLAT_CASES = ["splat_smooth", "splat_scatter", "splat_converge", "splat_border", "splat_identity"]
MODES = ["summation", "average", "linear", "softmax", "softmax_nometric"]


def _mods(cuda_lib):
    import fldr_vfi_b200.softSplat as S
    return S


def _run(S, x, fl, z, strType, gout=None):
    xd = x.cuda().requires_grad_(gout is not None)
    fd = fl.cuda().requires_grad_(gout is not None)
    zd = None if z is None else z.cuda().requires_grad_(gout is not None)
    y = S.FunctionSoftsplat(xd, fd, zd, strType)
    if gout is None:
        return y
    wrt = [xd, fd] + ([zd] if zd is not None and strType in ("linear", "softmax") else [])
    grads = torch.autograd.grad(y, wrt, gout.cuda())
    return y, grads


@pytest.mark.parametrize("name", SPLAT_CASES)
def test_raw_vs_golden(cuda_lib, name):
    S = _mods(cuda_lib)
    g = load_golden(name)
    xd = g["input"].cuda().requires_grad_(True)
    fd = g["flow"].cuda().requires_grad_(True)
    out = S._FunctionSoftsplat.apply(xd, fd)
    assert_splat_close(out, g["raw_out"], name + " raw fwd")
    gi, gf = torch.autograd.grad(out, [xd, fd], g["grad_out"].cuda())
    assert_splat_close(gi, g["raw_grad_input"], name + " raw gradInput")
    assert_splat_close(gf, g["raw_grad_flow"], name + " raw gradFlow")


@pytest.mark.parametrize("name", SPLAT_CASES)
@pytest.mark.parametrize("mode", MODES)
def test_wrapper_vs_golden(cuda_lib, name, mode):
    S = _mods(cuda_lib)
    g = load_golden(name)
    strType = "softmax" if mode == "softmax_nometric" else mode
    z = g["metric"] if mode in ("linear", "softmax") else None
    y, grads = _run(S, g["input"], g["flow"], z, strType, g["grad_out"])
    assert_splat_close(y, g[f"wrapper_{mode}_out"], f"{name} {mode} out", mag=None if mode in ("summation", "linear") else 1.0)
    assert_splat_close(grads[0], g[f"wrapper_{mode}_grad_input"], f"{name} {mode} grad_input")
    assert_splat_close(grads[1], g[f"wrapper_{mode}_grad_flow"], f"{name} {mode} grad_flow")
    if z is not None:
        assert_splat_close(grads[2], g[f"wrapper_{mode}_grad_metric"], f"{name} {mode} grad_metric")


@pytest.mark.parametrize("shape,regime,with_metric", [
    ((1, 3, 256, 448), "F1", True),      # BASELINE cfg1 image splat
    ((1, 16, 256, 448), "F1", False),    # cfg1 PCA features
    ((1, 48, 32, 56), "F2", False),      # true fLDR feature shape
    ((2, 3, 64, 96), "F3", True),        # max contention
    ((2, 5, 33, 47), "FB", True),        # ragged sizes, targets leaving the frame
    ((1, 1, 1, 1), "F0", True),          # degenerate
    ((3, 7, 5, 130), "F2", False),       # wide, C not multiple of 4
])
def test_softmax_vs_oracle_seeded(cuda_lib, shape, regime, with_metric):
    S = _mods(cuda_lib)
    N, C, H, W = shape
    x = synth.image(N, C, H, W, seed=11) if C == 3 else synth.features(N, C, H, W, seed=11)
    fl = synth.flow(N, H, W, regime, seed=12)
    z = synth.metric(N, H, W, seed=13) if with_metric else None
    gout = synth.grad(shape, seed=14)
    y, grads = _run(S, x, fl, z, "softmax", gout)
    yo, gi, gf, gz = so.function_softsplat_grads(x, fl, z, "softmax", gout)
    assert_splat_close(y, yo, "softmax out", mag=1.0)          # exactly the north_star bar: 1e-4 abs / 1e-5 rel
    assert_splat_close(grads[0], gi, "softmax grad_input")
    assert_splat_close(grads[1], gf, "softmax grad_flow")
    if z is not None:
        assert_splat_close(grads[2], gz, "softmax grad_metric")


def test_strided_views_and_expanded_metric(cuda_lib):
    """Callers pass views (fLDRnet.py:386-387,449-450): channel-sliced flow, frame-sliced 5-D input."""
    S = _mods(cuda_lib)
    N, H, W = 2, 40, 56
    x5 = synth.image(N, 3 * 2, H, W, seed=21).reshape(N, 3, 2, H, W)
    fl4 = torch.cat([synth.flow(N, H, W, "F1", seed=22) * 8, synth.flow(N, H, W, "F2", seed=23)], 1)
    z = synth.metric(N, H, W, seed=24)
    x5d, fl4d, zd = x5.cuda(), fl4.cuda(), z.cuda()
    for frame, sl in ((0, slice(0, 2)), (1, slice(2, 4))):
        y = S.FunctionSoftsplat(x5d[:, :, frame], fl4d[:, sl], zd, "softmax")
        yo = so.function_softsplat(x5[:, :, frame], fl4[:, sl], z, "softmax")
        assert y.is_contiguous()
        assert_splat_close(y, yo, f"view frame {frame}", mag=1.0)
    zc = torch.full((N, 1, 1, 1), -0.3)
    y = S.FunctionSoftsplat(x5d[:, :, 0], fl4d[:, :2], zc.cuda().expand(N, 1, H, W), "softmax")
    assert_splat_close(y, so.function_softsplat(x5[:, :, 0], fl4[:, :2], zc.expand(N, 1, H, W), "softmax"), "expanded metric", mag=1.0)


def test_needs_input_grad_and_no_grad(cuda_lib):
    S = _mods(cuda_lib)
    x = synth.image(1, 3, 16, 24, seed=31).cuda()
    fl = synth.flow(1, 16, 24, "F2", seed=32).cuda()
    z = synth.metric(1, 16, 24, seed=33).cuda()
    with torch.no_grad():
        y0 = S.Softsplat()(x, fl, z)
    assert not y0.requires_grad
    xr = x.clone().requires_grad_(True)
    y = S.Softsplat()(xr, fl.detach(), z)            # flow detached as at fLDRnet.py:384
    y.sum().backward()
    assert xr.grad is not None and fl.grad is None
    assert torch.equal(y.detach(), y0) or torch.allclose(y.detach(), y0, atol=1e-5)


def test_nonfinite_flow_is_skipped(cuda_lib):
    S = _mods(cuda_lib)
    x = synth.image(1, 3, 8, 8, seed=41)
    fl = torch.zeros(1, 2, 8, 8)
    fl[0, 0, 3, 3] = float("nan")
    fl[0, 1, 5, 5] = float("inf")
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), None, "softmax").cpu()
    ref = x.clone()
    ref[0, :, 3, 3] = -1
    ref[0, :, 5, 5] = -1
    assert torch.allclose(y, ref, atol=1e-6)


def test_nonfinite_flow_raises_the_debug_flag(cuda_lib):
    """The optional device word (fldr_splat_set_nonfinite_flag): set by a NaN / inf flow, untouched by a clean one; with
    FLDR_B200_CHECK_FLOW=1 the wrapper polls it and raises where the reference device-asserts (softSplat.py:25-26)."""
    import ctypes
    S = _mods(cuda_lib)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert cuda_lib.fldr_splat_set_nonfinite_flag(ctypes.c_void_p(flag.data_ptr())) == 0
    try:
        x = synth.image(1, 3, 64, 96, seed=41).cuda()
        fl = synth.flow(1, 64, 96, "F1", seed=42).cuda()
        for big in (False, True):                       # the single cooperative launch and the tile scatter kernel
            old = cuda_lib.fldr_get_option(b"splat_fused_max")
            cuda_lib.fldr_set_option(b"splat_fused_max", 0 if big else old)
            flag.zero_()
            S.FunctionSoftsplat(x, fl, None, "softmax")
            assert int(flag.item()) == 0
            bad = fl.clone()
            bad[0, 1, 5, 7] = float("inf")
            S.FunctionSoftsplat(x, bad, None, "softmax")
            assert int(flag.item()) == 1
            cuda_lib.fldr_set_option(b"splat_fused_max", old)
    finally:
        assert cuda_lib.fldr_splat_set_nonfinite_flag(None) == 0


def test_error_behaviour(cuda_lib):
    S = _mods(cuda_lib)
    x = torch.zeros(1, 3, 8, 8)
    fl = torch.zeros(1, 2, 8, 8)
    with pytest.raises(NotImplementedError):
        S.FunctionSoftsplat(x, fl, None, "softmax")                    # CPU tensors: softSplat.py:251-252
    with pytest.raises(AssertionError):
        S.FunctionSoftsplat(x.cuda(), fl.cuda(), None, "bogus")        # softSplat.py:322
    with pytest.raises(AssertionError):
        S.FunctionSoftsplat(x.cuda(), fl.cuda(), torch.zeros(1, 2, 8, 8).cuda(), "softmax")   # 321
    with pytest.raises(AssertionError):
        S.FunctionSoftsplat(x.cuda(), torch.zeros(1, 3, 8, 8).cuda(), None, "softmax")        # 227
    with pytest.raises(TypeError):
        S.FunctionSoftsplat(x.cuda().half(), fl.cuda(), None, "softmax")


# ------------------------------------------------------------------ full 4K size: properties, no oracle
H4K, W4K = 2304, 4096


def test_4k_identity_and_shift(cuda_lib):
    S = _mods(cuda_lib)
    x = synth.image(1, 3, H4K, W4K, seed=51).cuda()
    z = synth.metric(1, H4K, W4K, seed=52).cuda()
    fl = torch.zeros(1, 2, H4K, W4K, device="cuda")
    y = S.FunctionSoftsplat(x, fl, z, "softmax")
    assert float((y - x).abs().max()) <= 1e-6
    fl[:, 0] = 7.0
    fl[:, 1] = -5.0
    y = S.FunctionSoftsplat(x, fl, z, "softmax")
    assert float((y[:, :, :H4K - 5, 7:] - x[:, :, 5:, :W4K - 7]).abs().max()) <= 1e-6
    assert bool((y[:, :, H4K - 5:, :] == -1).all()) and bool((y[:, :, :, :7] == -1).all())


def test_4k_mass_conservation_and_linearity(cuda_lib):
    """Raw splat: each source spreads weights summing to 1, so with all targets in frame sum(S) == sum(in);
    and S is linear in its input."""
    S = _mods(cuda_lib)
    x1 = synth.image(1, 3, H4K, W4K, seed=53).cuda() + 2.0
    x2 = synth.image(1, 3, H4K, W4K, seed=54).cuda()
    fl = synth.flow(1, H4K, W4K, "F1", seed=55).cuda()
    # keep every target strictly inside the frame
    gx = torch.arange(W4K, device="cuda").view(1, 1, W4K)
    gy = torch.arange(H4K, device="cuda").view(1, H4K, 1)
    fl[:, 0] = (gx + fl[:, 0]).clamp(0, W4K - 1.001) - gx
    fl[:, 1] = (gy + fl[:, 1]).clamp(0, H4K - 1.001) - gy
    s1 = S._FunctionSoftsplat.apply(x1, fl)
    s2 = S._FunctionSoftsplat.apply(x2, fl)
    for c in range(3):
        tot_in = float(x1[:, c].double().sum())
        tot_out = float(s1[:, c].double().sum())
        assert abs(tot_in - tot_out) <= 1e-5 * abs(tot_in)
    s12 = S._FunctionSoftsplat.apply(0.5 * x1 + x2, fl)
    assert_splat_close(s12, 0.5 * s1 + s2, "4K linearity")


def test_4k_softmax_vs_oracle_full(cuda_lib):
    """cfg3 image splat (C=3 + metric, 2304x4096) against the oracle at full size (the oracle takes ~2 s)."""
    S = _mods(cuda_lib)
    x = synth.image(1, 3, H4K, W4K, seed=56)
    fl = synth.flow(1, H4K, W4K, "F1", seed=57)
    z = synth.metric(1, H4K, W4K, seed=58)
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), z.cuda(), "softmax")
    assert_splat_close(y, so.function_softsplat(x, fl, z, "softmax"), "4K softmax F1", mag=1.0)
    fl2 = synth.flow(1, H4K, W4K, "F2", seed=59)
    y2 = S.FunctionSoftsplat(x.cuda(), fl2.cuda(), z.cuda(), "softmax")
    assert_splat_close(y2, so.function_softsplat(x, fl2, z, "softmax"), "4K softmax F2", mag=1.0)


def test_feature_splat_L0_vs_oracle_full(cuda_lib):
    """cfg3 feature splat level 0 (C=48, metric=None, 288x512) forward + grad_input (flow detached, fLDRnet.py:384)."""
    S = _mods(cuda_lib)
    x = synth.features(1, 48, 288, 512, seed=71)
    fl = synth.flow(1, 288, 512, "F1", seed=72) * 8
    gout = synth.grad((1, 48, 288, 512), seed=73)
    xd = x.cuda().requires_grad_(True)
    y = S.Softsplat()(xd, fl.cuda())
    (gi,) = torch.autograd.grad(y, [xd], gout.cuda())
    yo, gio, _, _ = so.function_softsplat_grads(x, fl, None, "softmax", gout)
    assert_splat_close(y, yo, "feature L0 out")
    assert_splat_close(gi, gio, "feature L0 grad_input")


@pytest.fixture
def tile_everywhere(cuda_lib):
    """Send small frames through the three-pass path with the TMA tile scatter kernel too (they normally take the single
    cooperative launch)."""
    old = cuda_lib.fldr_get_option(b"splat_fused_max")
    cuda_lib.fldr_set_option(b"splat_fused_max", 0)
    yield
    cuda_lib.fldr_set_option(b"splat_fused_max", old)


@pytest.fixture
def plain_loads(cuda_lib):
    """Switch the TMA staging off: the plain-load scatter kernel (what odd views take) serves the call."""
    cuda_lib.fldr_set_option(b"splat_tma", 0)
    yield
    cuda_lib.fldr_set_option(b"splat_tma", 1)


@pytest.mark.parametrize("shape,regime,with_metric", [
    ((1, 3, 256, 448), "F1", True), ((1, 48, 32, 56), "F2", False), ((2, 5, 36, 48), "FB", False), ((3, 7, 8, 132), "F2", False),
    ((2, 3, 64, 96), "F3", True), ((1, 3, 4, 4), "F0", True), ((2, 3, 40, 260), "FB", True), ((1, 16, 37, 68), "F2", False),
])
def test_tile_kernel_small_frames(cuda_lib, tile_everywhere, shape, regime, with_metric):
    """Every quad shape / metric kind / ragged tile of the TMA tile scatter kernel against the oracle."""
    S = _mods(cuda_lib)
    N, C, H, W = shape
    x = synth.features(N, C, H, W, seed=11)
    fl = synth.flow(N, H, W, regime, seed=12)
    z = synth.metric(N, H, W, seed=13) if with_metric else None
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), None if z is None else z.cuda(), "softmax")
    assert_splat_close(y, so.function_softsplat(x, fl, z, "softmax"), "tile small", mag=1.0)


@pytest.mark.parametrize("mode", ["summation", "average", "linear"])
def test_tile_kernel_other_modes(cuda_lib, tile_everywhere, mode):
    S = _mods(cuda_lib)
    N, C, H, W = 2, 6 if mode != "linear" else 3, 40, 64
    x = synth.features(N, C, H, W, seed=21)
    fl = synth.flow(N, H, W, "F2", seed=22)
    z = synth.metric(N, H, W, seed=23) if mode == "linear" else None
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), None if z is None else z.cuda(), mode)
    assert_splat_close(y, so.function_softsplat(x, fl, z, mode), "tile " + mode)
    raw = S._FunctionSoftsplat.apply(x.cuda(), fl.cuda())
    assert_splat_close(raw, so.splat_raw(x, fl), "tile raw")


def test_plain_load_path_4k(cuda_lib, plain_loads):
    """The plain-load scatter kernel (odd views) at the 4K size."""
    S = _mods(cuda_lib)
    x = synth.image(1, 3, H4K, W4K, seed=56)
    fl = synth.flow(1, H4K, W4K, "F1", seed=57)
    z = synth.metric(1, H4K, W4K, seed=58)
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), z.cuda(), "softmax")
    assert_splat_close(y, so.function_softsplat(x, fl, z, "softmax"), "4K plain loads", mag=1.0)


def test_4k_converge_regime_vs_oracle(cuda_lib):
    """Flow regime F3 (every pixel flows to the frame centre: maximum contention, ~10^6 sources per target) at the full
    4K size.  With that many fp32 terms per pixel the summation order alone moves the result, so the bound adds 10x the
    oracle's own fp32-vs-fp64 discrepancy per element to the north_star tolerance."""
    S = _mods(cuda_lib)
    x = synth.image(1, 3, H4K, W4K, seed=56)
    fl = synth.flow(1, H4K, W4K, "F3", seed=60)
    z = synth.metric(1, H4K, W4K, seed=58)
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), z.cuda(), "softmax")
    ref32 = so.function_softsplat(x, fl, z, "softmax")
    ref64 = so.function_softsplat(x.double(), fl.double(), z.double(), "softmax")
    assert_splat_close(y, ref64, "4K softmax F3", mag=1.0, cond=(ref32.double() - ref64).abs() + 1e-5)
    assert float((y.cpu() == -1).float().mean()) > 0.99          # almost every pixel is a hole


def test_unaligned_views_take_plain_load_path(cuda_lib):
    """Views the TMA boxes cannot take (odd width, column-sliced rows, metric with >= 4 channels) run the plain-load kernel."""
    S = _mods(cuda_lib)
    N, H, W = 1, 300, 1027
    x = synth.image(N, 3, H, W, seed=61)
    fl = synth.flow(N, H, W, "F1", seed=62)
    z = synth.metric(N, H, W, seed=63)
    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), z.cuda(), "softmax")
    assert_splat_close(y, so.function_softsplat(x, fl, z, "softmax"), "odd width", mag=1.0)
    xs = x.cuda()[:, :, :, 3:1027]          # W = 1024 but rows start 12 bytes into a line
    fs, zs = fl.cuda()[:, :, :, 3:1027], z.cuda()[:, :, :, 3:1027]
    y = S.FunctionSoftsplat(xs, fs, zs, "softmax")
    assert_splat_close(y, so.function_softsplat(x[:, :, :, 3:1027], fl[:, :, :, 3:1027], z[:, :, :, 3:1027], "softmax"), "sliced rows", mag=1.0)
    x6 = synth.features(2, 6, 200, 256, seed=64)
    f6 = synth.flow(2, 200, 256, "F1", seed=65) * 4
    z6 = synth.metric(2, 200, 256, seed=66)
    y = S.FunctionSoftsplat(x6.cuda(), f6.cuda(), z6.cuda(), "softmax")
    assert_splat_close(y, so.function_softsplat(x6, f6, z6, "softmax"), "metric with 6 channels", mag=1.0)


def test_batched_tall_frames(cuda_lib):
    """N > 1 tall frames, forward and all gradients."""
    S = _mods(cuda_lib)
    N, H, W = 3, 640, 4096
    x = synth.image(N, 3, H, W, seed=91)
    z = synth.metric(N, H, W, seed=92)
    fl = synth.flow(N, H, W, "F1", seed=93)
    g = synth.grad((N, 3, H, W), seed=94)
    xd, fd, zd = x.cuda().requires_grad_(True), fl.cuda().requires_grad_(True), z.cuda().requires_grad_(True)
    y = S.FunctionSoftsplat(xd, fd, zd, "softmax")
    yo, gi, gf, gz = so.function_softsplat_grads(x, fl, z, "softmax", g)
    # fp64 oracle on the same fp32 coordinates' inputs: its distance from the fp32 oracle measures conditioning
    _, gi64, gf64, gz64 = so.function_softsplat_grads(x.double(), fl.double(), z.double(), "softmax", g.double())
    assert_splat_close(y, yo, "batched tall out", mag=1.0)
    grads = torch.autograd.grad(y, [xd, fd, zd], g.cuda())
    assert_splat_close(grads[0], gi, "batched tall grad_input", cond=gi - gi64)
    assert_splat_close(grads[1], gf, "batched tall grad_flow", cond=gf - gf64)
    assert_splat_close(grads[2], gz, "batched tall grad_metric", cond=gz - gz64)


def test_cfg5_training_shapes_vs_oracle(cuda_lib):
    """BASELINE configs[4] at its literal shapes: 32 x 3 x 512 x 512 image splat with all three gradients and the
    32 x 48 x 64 x 64 feature splat with grad_input (flow detached, fLDRnet.py:384)."""
    S = _mods(cuda_lib)
    N, H, W = 32, 512, 512
    x = synth.image(N, 3, H, W, seed=101)
    z = synth.metric(N, H, W, seed=102)
    fl = synth.flow(N, H, W, "F1", seed=103) * 4
    g = synth.grad((N, 3, H, W), seed=104)
    xd, fd, zd = x.cuda().requires_grad_(True), fl.cuda().requires_grad_(True), z.cuda().requires_grad_(True)
    y = S.FunctionSoftsplat(xd, fd, zd, "softmax")
    grads = torch.autograd.grad(y, [xd, fd, zd], g.cuda())
    yo, gi, gf, gz = so.function_softsplat_grads(x, fl, z, "softmax", g)
    _, gi64, gf64, gz64 = so.function_softsplat_grads(x.double(), fl.double(), z.double(), "softmax", g.double())
    assert_splat_close(y, yo, "cfg5 image out", mag=1.0)
    assert_splat_close(grads[0], gi, "cfg5 image grad_input", cond=gi - gi64)
    assert_splat_close(grads[1], gf, "cfg5 image grad_flow", cond=gf - gf64)
    assert_splat_close(grads[2], gz, "cfg5 image grad_metric", cond=gz - gz64)
    xf = synth.features(N, 48, 64, 64, seed=105)
    ff = synth.flow(N, 64, 64, "F1", seed=106) * 16
    gf_ = synth.grad((N, 48, 64, 64), seed=107)
    xfd = xf.cuda().requires_grad_(True)
    yf = S.Softsplat()(xfd, ff.cuda())
    (gif,) = torch.autograd.grad(yf, [xfd], gf_.cuda())
    yfo, gifo, _, _ = so.function_softsplat_grads(xf, ff, None, "softmax", gf_)
    assert_splat_close(yf, yfo, "cfg5 feature out")
    assert_splat_close(gif, gifo, "cfg5 feature grad_input")


@pytest.mark.parametrize("N,C,H,W,metric", [(1, 3, 96, 200, True), (3, 3, 40, 136, True), (2, 8, 33, 128, False)])
@pytest.mark.parametrize("regime", ["F1", "F2", "FB"])
def test_pass_order_and_prefetch_options_agree(cuda_lib, tile_everywhere, N, C, H, W, metric, regime):
    """The alternating row order of the three passes (splat_snake, default), the front-to-back order with cudaMemsetAsync and
    the L2 prefetch of accumulator rows are scheduling choices: every combination must give the oracle's result (the image-splat
    shape C = 3 + metric also exercises the cross / carry-to-E merges of the tile scatter under stretch, shear and border flows)."""
    import fldr_vfi_b200.softSplat as S
    x = (synth.image(N, C, H, W, seed=5) if C == 3 else synth.features(N, C, H, W, seed=5))
    fl = synth.flow(N, H, W, regime, seed=6) * (6.0 if regime == "F1" else 1.0)
    z = synth.metric(N, H, W, seed=7) if metric else None
    ref = so.function_softsplat(x, fl, z, "softmax")
    old = {k: cuda_lib.fldr_get_option(k) for k in (b"splat_snake", b"splat_pf_rows")}
    try:
        for snake in (1, 0):
            for pf in (0, 3, -1):
                cuda_lib.fldr_set_option(b"splat_snake", snake)
                cuda_lib.fldr_set_option(b"splat_pf_rows", pf)
                with torch.no_grad():
                    y = S.FunctionSoftsplat(x.cuda(), fl.cuda(), None if z is None else z.cuda(), "softmax")
                assert_splat_close(y, ref, f"snake={snake} pf={pf} {regime}")
    finally:
        for k, v in old.items():
            cuda_lib.fldr_set_option(k, v)

"""CPU tests: the oracle restatement against (i) the golden vectors produced by the reference's own kernel
text, (ii) that kernel text run live when oracle/_ref is available, (iii) analytic known-answer tests
(SURVEY.md App. A.4 - ours, the reference ships none)."""
import math

import pytest
import torch

from oracle import corr_oracle as co
from oracle import ref_host
from oracle import splat_oracle as so
from oracle import synth
from util import assert_corr_close, assert_splat_close, load_golden

SPLAT_CASES = ["splat_smooth", "splat_scatter", "splat_converge", "splat_border", "splat_identity"]
CORR_CASES = ["corr_c40", "corr_c32_odd", "corr_c196_tiny"]


@pytest.mark.parametrize("name", SPLAT_CASES)
def test_splat_raw_vs_golden(name):
    g = load_golden(name)
    assert_splat_close(so.splat_raw(g["input"], g["flow"]), g["raw_out"], name + " raw fwd")
    assert_splat_close(so.splat_raw_grad_input(g["flow"], g["grad_out"]), g["raw_grad_input"], name + " gradInput")
    assert_splat_close(so.splat_raw_grad_flow(g["input"], g["flow"], g["grad_out"]), g["raw_grad_flow"], name + " gradFlow")


@pytest.mark.parametrize("name", SPLAT_CASES)
@pytest.mark.parametrize("mode", ["summation", "average", "linear", "softmax", "softmax_nometric"])
def test_splat_wrapper_vs_golden(name, mode):
    g = load_golden(name)
    strType = "softmax" if mode == "softmax_nometric" else mode
    metric = g["metric"] if mode in ("linear", "softmax") else None
    y, gi, gf, gz = so.function_softsplat_grads(g["input"], g["flow"], metric, strType, g["grad_out"])
    assert_splat_close(y, g[f"wrapper_{mode}_out"], f"{name} {mode} out")
    assert_splat_close(gi, g[f"wrapper_{mode}_grad_input"], f"{name} {mode} grad_input")
    assert_splat_close(gf, g[f"wrapper_{mode}_grad_flow"], f"{name} {mode} grad_flow")
    if metric is not None:
        assert_splat_close(gz, g[f"wrapper_{mode}_grad_metric"], f"{name} {mode} grad_metric")


@pytest.mark.parametrize("name", CORR_CASES)
def test_corr_vs_golden(name):
    g = load_golden(name)
    f1, f2 = g["first"], g["second"]
    scale = co.correlation_fwd(f1.abs(), f2.abs())
    assert_corr_close(co.correlation_fwd(f1, f2), g["out"], scale, name + " fwd")
    gs = co.correlation_grad_first(f2.abs(), g["grad_out"].abs())
    assert_corr_close(co.correlation_grad_first(f2, g["grad_out"]), g["grad_first"], gs, name + " gradFirst")
    gs2 = co.correlation_grad_second(f1.abs(), g["grad_out"].abs())
    assert_corr_close(co.correlation_grad_second(f1, g["grad_out"]), g["grad_second"], gs2, name + " gradSecond")
    if "rbot0" in g:
        assert torch.equal(co.rearrange(f1), g["rbot0"])


@pytest.mark.skipif(not ref_host.available(), reason="oracle/_ref not built and /root/reference absent")
def test_oracle_vs_reference_kernel_text_live():
    """Fresh seeds (not in the fixtures): restatement vs the reference kernels executed on the host."""
    x = synth.features(1, 6, 19, 27, seed=123)
    fl = synth.flow(1, 19, 27, "F2", seed=124)
    gout = synth.grad((1, 6, 19, 27), seed=125)
    assert_splat_close(so.splat_raw(x, fl), ref_host.splat_update_output(x, fl), "live raw fwd")
    assert_splat_close(so.splat_raw_grad_input(fl, gout), ref_host.splat_update_grad_input(x, fl, gout), "live gradInput")
    assert_splat_close(so.splat_raw_grad_flow(x, fl, gout), ref_host.splat_update_grad_flow(x, fl, gout), "live gradFlow")
    f1 = synth.features(1, 35, 7, 10, seed=126)
    f2 = synth.features(1, 35, 7, 10, seed=127)
    assert_corr_close(co.correlation_fwd(f1, f2), ref_host.corr_update_output(f1, f2),
                      co.correlation_fwd(f1.abs(), f2.abs()), "live corr fwd")


def test_explicit_backward_matches_autograd_fp64():
    """The explicit gradInput / gradFlow restatements equal autograd of the fp64 forward (floor is constant)."""
    x = synth.features(1, 3, 10, 14, seed=5).double()
    fl = (synth.flow(1, 10, 14, "F2", seed=6)).double()
    g = synth.grad((1, 3, 10, 14), seed=7).double()
    xr = x.clone().requires_grad_(True)
    fr = fl.clone().requires_grad_(True)
    N, C, H, W = x.shape
    X, Y, x0, y0 = so._corners(fr)
    out = torch.zeros(N, C, H * W, dtype=torch.float64)
    for dx, dy, w in so._weights(X, Y, x0, y0):
        cx, cy = x0.long() + dx, y0.long() + dy
        valid = ((cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)).reshape(N, 1, H * W)
        idx = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).reshape(N, 1, H * W).expand(N, C, H * W)
        out = out.scatter_add(2, idx, torch.where(valid, xr.reshape(N, C, H * W) * w.reshape(N, 1, H * W), torch.zeros((), dtype=torch.float64)))
    gi, gf = torch.autograd.grad(out.reshape(N, C, H, W), [xr, fr], g)
    assert torch.allclose(gi, so.splat_raw_grad_input(fl, g), atol=1e-12)
    assert torch.allclose(gf, so.splat_raw_grad_flow(x, fl, g), atol=1e-12)


# ------------------------------------------------------------------ known-answer tests (App. A.4)
def test_kat_zero_flow_identity():
    x = synth.image(1, 3, 12, 16, seed=1)
    fl = torch.zeros(1, 2, 12, 16)
    z = synth.metric(1, 12, 16)
    assert torch.allclose(so.function_softsplat(x, fl, z, "softmax"), x, atol=1e-6)
    assert torch.allclose(so.function_softsplat(x, fl, None, "softmax"), x, atol=1e-6)
    assert torch.allclose(so.function_softsplat(x, fl, None, "summation"), (x - 0.5) * 2, atol=1e-6)


def test_kat_integer_shift_and_holes():
    x = synth.image(1, 3, 10, 12, seed=2)
    fl = torch.zeros(1, 2, 10, 12)
    fl[:, 0] = 3.0
    fl[:, 1] = -2.0
    y = so.function_softsplat(x, fl, None, "softmax")
    assert torch.allclose(y[:, :, :8, 3:], x[:, :, 2:, :9], atol=1e-6)
    assert bool((y[:, :, 8:, :] == -1).all()) and bool((y[:, :, :, :3] == -1).all())   # vacated pixels are holes -> -1


def test_kat_half_pixel_box_and_weighted_mean():
    x = synth.image(1, 1, 4, 8, seed=3)
    fl = torch.zeros(1, 2, 4, 8)
    fl[:, 0] = 0.5
    z = torch.full((1, 1, 4, 8), -0.7)
    y = so.function_softsplat(x, fl, z, "softmax")
    assert torch.allclose(y[:, :, :, 1:], 0.5 * (x[:, :, :, 1:] + x[:, :, :, :-1]), atol=1e-6)
    # two sources onto one target with z = (0, ln 3) -> 1:3 weighted mean
    x2 = torch.tensor([0.2, -0.6]).view(1, 1, 1, 2)
    f2 = torch.zeros(1, 2, 1, 2)
    f2[0, 0, 0, 1] = -1.0
    z2 = torch.tensor([0.0, math.log(3.0)]).view(1, 1, 1, 2)
    y2 = so.function_softsplat(x2, f2, z2, "softmax")
    assert abs(float(y2[0, 0, 0, 0]) - (0.25 * 0.2 + 0.75 * -0.6)) < 1e-6
    assert float(y2[0, 0, 0, 1]) == -1.0


def test_kat_correlation():
    f = synth.features(1, 8, 12, 14, seed=4)
    out = co.correlation_fwd(f, f)
    assert torch.allclose(out[:, 40], (f * f).mean(1), atol=1e-6)
    sh = torch.roll(f, shifts=(2, -3), dims=(2, 3))      # f2[y, x] = f[y-2, x+3]  -> f2[y+2, x-3] = f[y, x]
    out2 = co.correlation_fwd(f, sh)
    ch = (2 + 4) * 9 + (-3 + 4)
    assert torch.allclose(out2[:, ch, 4:-4, 4:-4], (f * f).mean(1)[:, 4:-4, 4:-4], atol=1e-6)
    # the 4-pixel border ring sees zero padding: displacement (-4,-4) at pixel (0,0) reads outside the frame
    assert float(out[0, 0, 0, 0]) == 0.0


# ---------------------------------------------------------------- backward warp + splat metric
WARP_CASES = ["warp_smooth", "warp_scatter", "warp_border", "warp_identity"]


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_oracle_vs_golden(name):
    """The explicit restatement against the outputs of the reference's own bwarp source (tests/golden/make_golden.py)."""
    from oracle import warp_oracle as wo
    g = load_golden(name)
    a = float(g["alpha"])
    for key, got in (("bwarp_mask", wo.bwarp(g["src"], g["flow"], True)), ("bwarp_nomask", wo.bwarp(g["src"], g["flow"], False)),
                     ("metric", wo.warp_metric(g["ref"], g["src"], g["flow"], a))):
        err = float((got - g[key]).abs().max())
        assert err <= 2e-6, (name, key, err)
    # the mask decision itself is identical: no pixel kept on one side and zeroed on the other
    kept_ref = g["bwarp_mask"].abs().sum(1) > 0
    kept_got = wo.bwarp(g["src"], g["flow"], True).abs().sum(1) > 0
    assert bool((kept_ref == kept_got).all())


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_oracle_gradients_vs_golden(name):
    """Autograd through the explicit restatement against autograd through the reference's own bwarp (grid_sample
    backward), both on the CPU."""
    from oracle import warp_oracle as wo
    g = load_golden(name)
    xi, fi = g["src"].clone().requires_grad_(True), g["flow"].clone().requires_grad_(True)
    gx, gf = torch.autograd.grad(wo.bwarp(xi, fi, True), [xi, fi], g["grad_out"])
    assert float((gx - g["grad_src"]).abs().max()) <= 2e-6 * max(1.0, float(g["grad_src"].abs().max()))
    assert float((gf - g["grad_flow"]).abs().max()) <= 2e-6 * max(1.0, float(g["grad_flow"].abs().max()))


PWCWARP_CASES = ["pwcwarp_smooth", "pwcwarp_scatter", "pwcwarp_border"]


@pytest.mark.parametrize("name", PWCWARP_CASES)
def test_pwc_backward_oracle_vs_golden(name):
    """PWC-Net's Backward restated, against the outputs of its own source (OpticalFlow/PWCNet.py:116-143)."""
    from oracle import warp_oracle as wo
    g = load_golden(name)
    got = wo.pwc_backward(g["input"], g["flow"])
    assert float((got - g["out"]).abs().max()) <= 1e-6 * max(1.0, float(g["out"].abs().max()))
    assert bool(((got.abs().sum(1) > 0) == (g["out"].abs().sum(1) > 0)).all())


def test_linspace_restatement_is_bit_exact():
    from oracle import warp_oracle as wo
    for n in (1, 2, 3, 17, 23, 40, 64, 288, 1024, 2304, 4096):
        assert bool((wo._linspace_pm1(n) == torch.linspace(-1.0, 1.0, n)).all()), n


def test_warp_kat_zero_flow_is_not_identity_but_integer_grid_is():
    """Reference quirk (fLDRnet.py:565-568): coordinates are normalised with W-1 and sampled with align_corners=False,
    so zero flow resamples at X*W/(W-1) - 0.5.  A flow of u = (x + 0.5)*(W-1)/W - x undoes it exactly."""
    from oracle import warp_oracle as wo
    H, W = 6, 8
    x = synth.image(1, 3, H, W, seed=7)
    zero = wo.bwarp(x, torch.zeros(1, 2, H, W), withmask=False)
    assert float((zero - x).abs().max()) > 1e-3
    gx = torch.arange(W, dtype=torch.float64).view(1, 1, 1, W).expand(1, 1, H, W)
    gy = torch.arange(H, dtype=torch.float64).view(1, 1, H, 1).expand(1, 1, H, W)
    fl = torch.cat([(gx + 0.5) * (W - 1) / W - gx, (gy + 0.5) * (H - 1) / H - gy], 1).float()
    ident, msum = wo.bwarp(x, fl, withmask=True, return_mask=True)
    assert float((ident - x).abs().max()) < 1e-5 and float(msum.min()) > 0.999


def test_warp_kat_mask_and_metric():
    from oracle import warp_oracle as wo
    H, W = 8, 8
    x = torch.ones(1, 2, H, W)
    fl = torch.zeros(1, 2, H, W)
    fl[:, 0] = 100.0                                              # everything samples outside the frame
    assert float(wo.bwarp(x, fl).abs().max()) == 0.0
    z = wo.warp_metric(x, x, fl, -2.0)                            # |1 - 0| * -2, mean over channels
    assert z.shape == (1, 1, H, W) and float((z + 2.0).abs().max()) == 0.0


# ---------------------------------------------------------------- occlusion softmax + image synthesis
BLEND_CASES = ["blend_t05", "blend_temp", "blend_c1"]


@pytest.mark.parametrize("name", BLEND_CASES)
def test_blend_oracle_vs_golden(name):
    """The restatement against the outputs of fLDRnet.py's own lines 510-524 (tests/golden/make_golden.py)."""
    from oracle import blend_oracle as bo
    g = load_golden(name)
    T = g["temperature"].reshape(1).double()
    out, occ0 = bo.occ_blend(g["refine_out"], T, g["t_value"], *[g[f"img{k}"] for k in range(6)])
    assert out.dtype == torch.float64 and g["out"].dtype == torch.float64
    assert float((out - g["out"]).abs().max()) <= 1e-15 and float((occ0 - g["occ0"]).abs().max()) <= 1e-15


def test_blend_kat():
    """Equal logits and t = 0.5: every weight is 1/12, so the result is the plain mean of the six images; one dominant
    logit selects its image whatever t is (the divisor normalises t away)."""
    from oracle import blend_oracle as bo
    N, C, H, W = 1, 3, 4, 5
    imgs = [synth.image(N, C, H, W, seed=300 + k) for k in range(6)]
    T = torch.ones(1, dtype=torch.float64)
    out, occ0 = bo.occ_blend(torch.zeros(N, 6, H, W), T, torch.full((N, 1), 0.5), *imgs)
    assert float((out - sum(i.double() for i in imgs) / 6).abs().max()) < 1e-15
    assert float((occ0 - 1 / 6).abs().max()) < 1e-15
    logits = torch.zeros(N, 6, H, W)
    logits[:, 3] = 800.0
    out, _ = bo.occ_blend(logits, T, torch.full((N, 1), 0.3), *imgs)
    assert float((out - imgs[3].double()).abs().max()) < 1e-12


# ------------------------------------------------------------------ block-PCA features
@pytest.mark.parametrize("name", ["pca_6x32x48", "pca_12x16x24_nomv", "pca_6x8x8"])
def test_pca_restatement_vs_reference_text(name):
    """oracle/pca_oracle.py against the golden vectors produced by the reference's own to_pca_diff text (float64)."""
    from oracle import pca_oracle
    g = load_golden(name)
    mv = g["mean_vec"] if int(g["mean_vector_norm"]) else None
    out = pca_oracle.to_pca_diff(g["im"], g["mean"], g["EV"], mv)
    assert out.dtype == torch.float64 and out.shape == g["out"].shape
    assert float((out - g["out"]).abs().max()) <= 1e-13
    assert float(out.min()) == -1.0 and float(out.max()) == 1.0


def test_pca_rejects_unpadded_frames():
    from oracle import pca_oracle
    with pytest.raises(Exception):
        pca_oracle.to_pca_diff(torch.zeros(6, 12, 16), torch.zeros(64, dtype=torch.float64), torch.zeros(16, 64, dtype=torch.float64))


PYRAMID_CASES = ["pyramid_5lv", "pyramid_noise_b2", "pyramid_ac", "pyramid_odd"]
PYRAMID_TOL = 2e-6       # float32, 16 taps: x max|frame| - fma contraction / vector order of ATen's kernel vs the plain restatement


@pytest.mark.parametrize("name", PYRAMID_CASES)
def test_pyramid_restatement_vs_reference_expression(name):
    """oracle/pyramid_oracle.py against the reference's own list comprehension (main.py:855-856) evaluated on the CPU."""
    import numpy as np
    from oracle import pyramid_oracle as po
    g = load_golden(name)
    frames = g["frames"].numpy()
    n = int(g["n_levels"])
    levels = po.input_pyramid(frames, [int(s) for s in g["scales"]], n, bool(int(g["align_corners"])))
    assert len(levels) == n + 1 and levels[0].shape == frames.shape
    for i in range(1, n + 1):
        ref = g[f"level{i}"].numpy()
        assert levels[i].shape == ref.shape and levels[i].dtype == np.float32
        assert float(np.abs(levels[i] - ref).max()) <= PYRAMID_TOL * max(1.0, float(np.abs(frames).max()))


def test_pyramid_kat_constant_and_half_pixel_weights():
    """Known answers: a constant frame stays constant on every level (the cubic weights sum to 1), and for the factors
    1/2^k every output uses the weights (-3/32, 19/32, 19/32, -3/32) at rows / columns 2^k d + 2^(k-1) - 2 .. + 1."""
    import numpy as np
    from oracle import pyramid_oracle as po
    const = np.full((1, 16, 32), 0.37, dtype=np.float32)
    for f in (0.5, 0.25, 1 / 3):
        assert float(np.abs(po.bicubic_resize(const, f) - np.float32(0.37)).max()) <= 1e-7
    for k in (1, 2, 3):
        taps, w = po.axis_taps(64, 64 >> k, 1.0 / (1 << k), False)
        assert np.array_equal(w, np.tile(np.float32([-0.09375, 0.59375, 0.59375, -0.09375]), (64 >> k, 1)))
        d = np.arange(64 >> k)[:, None]
        assert np.array_equal(taps, np.clip((d << k) + (1 << (k - 1)) - 2 + np.arange(4)[None, :], 0, 63))

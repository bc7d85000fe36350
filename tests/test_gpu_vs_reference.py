"""GPU tests against the UNMODIFIED reference op files running their own CuPy kernels on the same GPU (through
baseline/cupy_shim + NVRTC).  Skipped when baseline/_ref was not staged (python baseline/fetch_ref.py in the build
container).  This is the second pin of the parity claim: reference code, reference kernels, same device."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_gpu  # noqa: E402
from oracle import corr_oracle as co  # noqa: E402
from oracle import synth  # noqa: E402
from util import assert_corr_close, assert_splat_close  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_gpu.available(), reason="baseline/_ref not staged")]


@pytest.mark.parametrize("strType,with_metric", [("softmax", True), ("softmax", False), ("average", False), ("linear", True), ("summation", False)])
def test_splat_matches_reference_kernels(cuda_lib, strType, with_metric):
    import fldr_vfi_b200.softSplat as S
    R = ref_gpu.softsplat_module()
    N, C, H, W = 2, 3, 96, 160
    x = synth.image(N, C, H, W, seed=1).cuda()
    fl = (synth.flow(N, H, W, "F1", seed=2) * 10).cuda()
    z = synth.metric(N, H, W, seed=3).cuda() if with_metric else None
    g = synth.grad((N, C, H, W), seed=4).cuda()
    outs = []
    for mod in (S, R):
        xd, fd = x.clone().requires_grad_(True), fl.clone().requires_grad_(True)
        zd = None if z is None else z.clone().requires_grad_(True)
        y = mod.FunctionSoftsplat(xd, fd, zd, strType)
        wrt = [xd, fd] + ([zd] if zd is not None else [])
        outs.append((y.detach(), torch.autograd.grad(y, wrt, g)))
    (y, gr), (yr, grr) = outs
    assert_splat_close(y, yr, f"{strType} out vs reference kernels", mag=None if strType in ("summation", "linear") else 1.0)
    for a, b, nm in zip(gr, grr, ("grad_input", "grad_flow", "grad_metric")):
        assert_splat_close(a, b, f"{strType} {nm} vs reference kernels")


def test_raw_splat_matches_reference_kernels(cuda_lib):
    import fldr_vfi_b200.softSplat as S
    R = ref_gpu.softsplat_module()
    x = synth.features(1, 6, 64, 80, seed=5).cuda()
    fl = synth.flow(1, 64, 80, "F2", seed=6).cuda()
    assert_splat_close(S._FunctionSoftsplat.apply(x, fl), R._FunctionSoftsplat.apply(x, fl), "raw vs reference kernels")


def test_4k_image_splat_matches_reference_kernels(cuda_lib):
    import fldr_vfi_b200.softSplat as S
    R = ref_gpu.softsplat_module()
    x = synth.image(1, 3, 2304, 4096, seed=56).cuda()
    fl = synth.flow(1, 2304, 4096, "F1", seed=57).cuda()
    z = synth.metric(1, 2304, 4096, seed=58).cuda()
    with torch.no_grad():
        assert_splat_close(S.Softsplat()(x, fl, z), R.Softsplat()(x, fl, z), "4K image splat vs reference kernels", mag=1.0)


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 80, 128), (2, 196, 5, 8), (1, 40, 24, 36)])
def test_correlation_matches_reference_kernels(cuda_lib, B, C, H, W):
    import fldr_vfi_b200.correlation as Cm
    R = ref_gpu.correlation_module()
    f1 = synth.features(B, C, H, W, seed=3)
    f2 = synth.features(B, C, H, W, seed=5)
    g = synth.grad((B, 81, H, W), seed=4)
    res = []
    for mod in (Cm, R):
        a, b = f1.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
        out = mod.FunctionCorrelation(tensorFirst=a, tensorSecond=b)
        res.append((out.detach(), torch.autograd.grad(out, [a, b], g.cuda())))
    (o, gr), (orf, grr) = res
    assert_corr_close(o, orf, co.correlation_fwd(f1.abs(), f2.abs()), "corr fwd vs reference kernels")
    assert_corr_close(gr[0], grr[0], co.correlation_grad_first(f2.abs(), g.abs()), "corr gradFirst vs reference kernels")
    assert_corr_close(gr[1], grr[1], co.correlation_grad_second(f1.abs(), g.abs()), "corr gradSecond vs reference kernels")

"""GPU parity tests for the occlusion softmax + image synthesis row: CUDA path through the C-ABI vs
the float64 oracle and the golden vectors of fLDRnet.py's own lines 510-524; views, t per sample, the 4K shape against the
reference's own statements (lifted from baseline/_ref) run on the same GPU."""
import pytest
import torch

from oracle import blend_oracle as bo
from oracle import synth
from util import load_golden

pytestmark = pytest.mark.gpu

BLEND_CASES = ["blend_t05", "blend_temp", "blend_c1"]
TOL = 2e-14      # float64 end to end; only exp() may differ from the host libm by an ulp


def _mod(cuda_lib):
    import fldr_vfi_b200.blend as Bl
    return Bl


@pytest.mark.parametrize("name", BLEND_CASES)
def test_vs_golden(cuda_lib, name):
    Bl = _mod(cuda_lib)
    g = load_golden(name)
    T = g["temperature"].reshape(1).double().cuda()
    with torch.no_grad():
        out, occ0 = Bl.occ_blend(g["refine_out"].cuda(), T, g["t_value"].cuda(), *[g[f"img{k}"].cuda() for k in range(6)],
                                 return_occ0=True)
    assert out.dtype == torch.float64 and out.is_contiguous() and occ0.shape == g["occ0"].shape
    assert float((out.cpu() - g["out"]).abs().max()) <= TOL, name
    assert float((occ0.cpu() - g["occ0"]).abs().max()) <= TOL, name


@pytest.mark.parametrize("N,C,H,W", [(1, 3, 37, 61), (3, 3, 16, 130), (2, 5, 9, 7), (1, 1, 1, 1)])
def test_vs_oracle_shapes_and_views(cuda_lib, N, C, H, W):
    Bl = _mod(cuda_lib)
    imgs = [synth.image(N, C, H, W, seed=400 + k) for k in range(4)]
    x_l = torch.stack([synth.image(N, C, H, W, seed=410), synth.image(N, C, H, W, seed=411)], 2)     # [N,C,2,H,W]: strided x0, x1
    refine = synth.grad((N, 9, H, W), seed=420) * 4.0
    t = torch.rand(N, 1, generator=torch.Generator().manual_seed(5))
    T = torch.tensor([0.73], dtype=torch.float64)
    want, occ_w = bo.occ_blend(refine, T, t, *imgs, x_l[:, :, 0], x_l[:, :, 1])
    xd = x_l.cuda()
    with torch.no_grad():
        got, occ_g = Bl.occ_blend(refine.cuda(), T.cuda(), t.cuda().view(N, 1, 1, 1), *[i.cuda() for i in imgs], xd[:, :, 0], xd[:, :, 1],
                                  return_occ0=True)
    assert float((got.cpu() - want).abs().max()) <= TOL
    assert float((occ_g.cpu() - occ_w).abs().max()) <= TOL


def test_errors(cuda_lib):
    Bl = _mod(cuda_lib)
    x = torch.zeros(1, 3, 8, 8, device="cuda")
    lg = torch.zeros(1, 6, 8, 8, device="cuda")
    t = torch.full((1, 1), 0.5, device="cuda")
    with pytest.raises(TypeError):                      # a float32 temperature would silently change the arithmetic type
        Bl.occ_blend(lg, torch.ones(1, device="cuda"), t, x, x, x, x, x, x)
    with pytest.raises(NotImplementedError):             # occ_0 is a forward-only extra
        Bl.occ_blend(lg.requires_grad_(True), torch.ones(1, dtype=torch.float64, device="cuda"), t, x, x, x, x, x, x, return_occ0=True)


def test_gradients_vs_reference_lines_under_autograd(cuda_lib):
    """Backward of the blend (training differentiates through fLDRnet.py:509-524: the logits, the splatted images, and T_param
    when TOptimization is on) against autograd through the oracle's restatement of those lines."""
    from oracle import blend_oracle as bo
    from oracle import synth
    Bl = _mod(cuda_lib)
    N, C, H, W = 2, 3, 12, 20
    imgs = [synth.image(N, C, H, W, seed=300 + k) for k in range(6)]
    lg = synth.grad((N, 8, H, W), seed=310) * 2.0
    t = torch.tensor([[0.5], [0.3]])
    T = torch.full((1,), 0.8, dtype=torch.float64)
    go = synth.grad((N, C, H, W), seed=311).double()
    lr, Tr = lg.clone().requires_grad_(True), T.clone().requires_grad_(True)
    ir = [im.clone().requires_grad_(True) for im in imgs]
    oo, _ = bo.occ_blend(lr, Tr, t, *ir)
    ref = torch.autograd.grad(oo, [lr, Tr] + ir, go)
    ld, Td = lg.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
    idv = [im.cuda().requires_grad_(True) for im in imgs]
    out = Bl.occ_blend(ld, Td, t.cuda(), *idv)
    assert float((out.detach().cpu() - oo.detach()).abs().max()) <= 2e-14
    got = torch.autograd.grad(out, [ld, Td] + idv, go.cuda())
    for a, b, nm in zip(got, ref, ["logits", "T_param"] + [f"image{k}" for k in range(6)]):
        assert a.dtype == b.dtype, nm
        assert float((a.cpu().double() - b.double()).abs().max()) <= 1e-6 * max(1.0, float(b.abs().max())), nm


def test_4k_vs_reference_operator_sequence(cuda_lib):
    Bl = _mod(cuda_lib)
    H, W = 2304, 4096
    imgs = [synth.image(1, 3, H, W, seed=500 + k).cuda() for k in range(6)]
    refine = (synth.grad((1, 6, H, W), seed=510) * 3.0).cuda()
    t_value = torch.full((1, 1, 1, 1), 0.5, device="cuda")
    T = torch.ones(1, dtype=torch.float64, device="cuda")
    from baseline import ref_src
    if not ref_src.available():
        pytest.skip("baseline/_ref not staged")
    with torch.no_grad():
        got = Bl.occ_blend(refine, T, t_value, *imgs)
        out, _ = ref_src.blend()(refine, T, t_value, imgs[0], imgs[1], imgs[2], imgs[3], torch.stack([imgs[4], imgs[5]], 2))
    assert out.dtype == torch.float64
    assert float((got - out).abs().max()) <= 1e-13
    # convexity: the blend of images in [-1, 1] stays in [-1, 1]
    assert float(got.abs().max()) <= 1.0 + 1e-12

"""GPU parity tests of the block-PCA feature extraction (pca_comp.py:473-528) through the C-ABI: golden
vectors from the reference's own function text, the oracle on seeded shapes incl. the 4K frame pair, error behaviour."""
import types

import pytest
import torch

from oracle import pca_oracle, synth
from util import load_golden

pytestmark = pytest.mark.gpu
TOL = 2e-13          # float64, sequential fma per output vs the reference's DGEMM: a few ulp of O(1) values


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    mean = torch.randn(64, generator=g, dtype=torch.float64) * 0.1
    EV = torch.linalg.qr(torch.randn(64, 64, generator=g, dtype=torch.float64))[0][:16].contiguous()
    mean_vec = torch.rand(16, generator=g, dtype=torch.float64) + 0.5
    return mean, EV, mean_vec


@pytest.mark.parametrize("name", ["pca_6x32x48", "pca_12x16x24_nomv", "pca_6x8x8"])
def test_pca_vs_golden(cuda_lib, name):
    import fldr_vfi_b200.pca as P
    g = load_golden(name)
    mv = g["mean_vec"].cuda() if int(g["mean_vector_norm"]) else None
    with torch.no_grad():
        out = P.pca_features(g["im"].cuda(), g["mean"].cuda(), g["EV"].cuda(), mv)
    assert out.dtype == torch.float64 and tuple(out.shape) == tuple(g["out"].shape)
    assert float((out.cpu() - g["out"]).abs().max()) <= TOL


def test_to_pca_diff_signature_matches_reference(cuda_lib):
    """Same call as fLDRnet.py:146 makes: to_pca_diff(im, params, args, mean, EV, mean_vec)."""
    import fldr_vfi_b200.pca as P
    g = load_golden("pca_6x32x48")
    params = types.SimpleNamespace(wiS=8, weightMat=None, components_fraction=0.25)
    args = types.SimpleNamespace(gpu=0, mean_vector_norm=True)
    with torch.no_grad():
        out = P.to_pca_diff(g["im"].cuda(), params, args, g["mean"].cuda(), g["EV"].cuda(), g["mean_vec"].cuda())
    assert float((out.cpu() - g["out"]).abs().max()) <= TOL


@pytest.mark.parametrize("chan,H,W,use_mv", [(6, 64, 136, True), (12, 8, 2048, False), (6, 2304, 4096, True)])
def test_pca_vs_oracle_seeded(cuda_lib, chan, H, W, use_mv):
    """Ragged block columns (136 / 8 = 17 blocks: a partial CTA), wide frames, and the literal 4K shape of cfg3."""
    import fldr_vfi_b200.pca as P
    im = synth.image(1, chan, H, W, seed=7)[0]
    mean, EV, mv = _params(8)
    ref = pca_oracle.to_pca_diff(im, mean, EV, mv if use_mv else None)
    with torch.no_grad():
        out = P.pca_features(im.cuda(), mean.cuda(), EV.cuda(), mv.cuda() if use_mv else None)
        out32 = P.pca_features(im.cuda(), mean.cuda(), EV.cuda(), mv.cuda() if use_mv else None, out_dtype=torch.float32)
    assert float((out.cpu() - ref).abs().max()) <= TOL
    assert out32.dtype == torch.float32 and torch.equal(out32.cpu(), ref.float()) or float((out32.cpu() - ref.float()).abs().max()) <= 6e-8
    assert float(out.min()) == -1.0 and float(out.max()) == 1.0


def test_pca_strided_views_and_errors(cuda_lib):
    import fldr_vfi_b200.pca as P
    mean, EV, mv = _params(9)
    x5 = synth.image(2, 6, 32, 48, seed=10).reshape(2, 3, 2, 32, 48)        # x_l[i]: [B, 3, 2, H, W]
    im = x5.reshape(12, 32, 48)
    ref = pca_oracle.to_pca_diff(im, mean, EV, mv)
    with torch.no_grad():
        out = P.pca_features(x5.cuda().reshape(12, 32, 48), mean.cuda(), EV.cuda(), mv.cuda())
        wide = torch.zeros(12, 32, 64).cuda()
        wide[:, :, :48] = im.cuda()
        out_v = P.pca_features(wide[:, :, :48], mean.cuda(), EV.cuda(), mv.cuda())      # row stride 64, width 48
    assert float((out.cpu() - ref).abs().max()) <= TOL and float((out_v.cpu() - ref).abs().max()) <= TOL
    with pytest.raises(Exception):
        P.pca_features(torch.zeros(6, 12, 16).cuda(), mean.cuda(), EV.cuda())                # not padded: pca_comp.py:486-487
    with pytest.raises(NotImplementedError):
        P.pca_features(im, mean, EV)                                                         # CPU tensors
    with pytest.raises(TypeError):
        P.pca_features(im.cuda(), mean.float().cuda(), EV.cuda())
    with pytest.raises(NotImplementedError):
        P.pca_features(im.cuda(), mean.cuda(), EV.cuda().requires_grad_(True))

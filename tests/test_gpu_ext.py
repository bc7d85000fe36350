"""The two host bindings of the same C ABI - the thin torch C++ extension (``_fldr_torch_ext``, what the ops use when it is
built) and the ctypes table (``_lib.SYMBOLS``, what any foreign-language consumer would write) - give the same results."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _run_all(S, C):
    x = synth.image(2, 3, 40, 64, seed=1).cuda().requires_grad_(True)
    fl = (synth.flow(2, 40, 64, "F1", seed=2) * 6).cuda().requires_grad_(True)
    z = synth.metric(2, 40, 64, seed=3).cuda().requires_grad_(True)
    g = synth.grad((2, 3, 40, 64), seed=4).cuda()
    y = S.FunctionSoftsplat(x, fl, z, "softmax")
    gs = torch.autograd.grad(y, [x, fl, z], g)
    raw = S._FunctionSoftsplat.apply(x, fl)
    gr = torch.autograd.grad(raw, [x, fl], g)
    a = synth.features(2, 24, 16, 24, seed=5).cuda().requires_grad_(True)
    b = synth.features(2, 24, 16, 24, seed=6).cuda().requires_grad_(True)
    o = C.FunctionCorrelation(tensorFirst=a, tensorSecond=b)
    gc = torch.autograd.grad(o, [a, b], synth.grad((2, 81, 16, 24), seed=7).cuda())
    return [y.detach(), *gs, raw.detach(), *gr, o.detach(), *gc]


def test_extension_and_ctypes_bindings_agree(cuda_lib):
    import fldr_vfi_b200._lib as L
    import fldr_vfi_b200.correlation as C
    import fldr_vfi_b200.softSplat as S
    ext = L.ext()
    if ext is None:
        pytest.skip("_fldr_torch_ext not built (python fldr-vfi_b200/build_ext.py): the ctypes binding serves every call")
    assert ext.abi_version() == cuda_lib.fldr_abi_version()
    via_ext = _run_all(S, C)
    saved = L._ext
    L._ext = None                       # force the ctypes binding
    try:
        via_ctypes = _run_all(S, C)
    finally:
        L._ext = saved
    for i, (p, q) in enumerate(zip(via_ext, via_ctypes)):
        assert p.shape == q.shape and p.dtype == q.dtype, i
        assert float((p - q).abs().max()) <= 1e-5 * max(1.0, float(q.abs().max())), i       # atomics: summation order only

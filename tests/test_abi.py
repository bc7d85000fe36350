"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/fldr_b200.h declares; argument validation that needs no device; host-mirror error behaviour."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, os.path.join(ROOT, "fldr-vfi_b200"))
    import build as fldr_build
    sys.path.pop(0)
    fldr_build.build(verbose=False)
    import fldr_vfi_b200._lib as L
    return L.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fldr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fldr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    import fldr_vfi_b200._lib as L
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fldr_b200.h but not exported"
    assert sorted(L.SYMBOLS) == names, "ctypes binding table and header disagree"


def test_status_strings_and_version(lib):
    assert lib.fldr_abi_version() == 1
    assert lib.fldr_status_string(0) == b"FLDR_OK"
    for code in (-1, -2, -3, -4, -5):
        assert lib.fldr_status_string(code).startswith(b"FLDR_ERR_")
    assert lib.fldr_status_string(-99) == b"FLDR_ERR_UNKNOWN"


def test_options_roundtrip(lib):
    """fldr_set_option / fldr_get_option: every documented switch exists, unknown names are rejected."""
    for name in (b"splat_tma", b"splat_fused_max", b"corr_th", b"splat_pf_rows", b"corr_bwd_rows", b"splat_snake"):
        old = lib.fldr_get_option(name)
        assert lib.fldr_set_option(name, 7) == 0 and lib.fldr_get_option(name) == 7
        assert lib.fldr_set_option(name, old) == 0
    assert lib.fldr_set_option(b"no_such_option", 1) == -1
    assert lib.fldr_set_option(None, 1) == -1
    assert lib.fldr_get_option(b"no_such_option") == 0


def test_workspace_sizes(lib):
    # 4K image splat: the accumulator, rows of W + 2 float4 cells (one guard cell either side)
    acc = 2304 * 4098 * 16
    ws = lib.fldr_splat_fwd_workspace_bytes(3, 1, 3, 2304, 4096)
    assert acc <= ws <= acc + 256
    # feature splat level 0: 13 channel quads
    ws = lib.fldr_splat_fwd_workspace_bytes(3, 1, 48, 288, 512)
    assert 13 * 288 * 514 * 16 <= ws <= 13 * 288 * 514 * 16 + 256
    assert lib.fldr_splat_fwd_workspace_bytes(9, 1, 3, 8, 8) == 0          # unknown mode
    # the size depends on the shape only: tuning options must not change it (a size cached per shape stays valid)
    ws0 = lib.fldr_splat_fwd_workspace_bytes(3, 1, 3, 2304, 4096)
    for name, val in ((b"splat_tma", 0), (b"splat_fused_max", 0), (b"splat_pf_rows", 8), (b"splat_snake", 0)):
        old = lib.fldr_get_option(name)
        lib.fldr_set_option(name, val)
        assert lib.fldr_splat_fwd_workspace_bytes(3, 1, 3, 2304, 4096) == ws0
        lib.fldr_set_option(name, old)


def test_argument_validation_needs_no_device(lib):
    s4 = (ctypes.c_int64 * 4)(0, 0, 0, 0)
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    # null input
    assert lib.fldr_splat_fwd(3, None, s4, p, s4, None, None, p, None, 1, 3, 8, 8, p, 1 << 20, None) == -1
    # unknown mode / bad sizes
    assert lib.fldr_splat_fwd(7, p, s4, p, s4, None, None, p, None, 1, 3, 8, 8, p, 1 << 20, None) == -1
    assert lib.fldr_splat_fwd(3, p, s4, p, s4, None, None, p, None, 1, 0, 8, 8, p, 1 << 20, None) == -1
    # linear without metric (softSplat.py:328)
    assert lib.fldr_splat_fwd(2, p, s4, p, s4, None, None, p, None, 1, 3, 8, 8, p, 1 << 20, None) == -4
    # workspace too small
    assert lib.fldr_splat_fwd(3, p, s4, p, s4, None, None, p, None, 1, 3, 8, 8, p, 16, None) == -2
    # maximum sizes: planes of 2^31 pixels and more are refused up front, not mis-indexed (the reference's int32
    # element index silently overflows there, softSplat.py:18-22)
    assert lib.fldr_splat_fwd(3, p, s4, p, s4, None, None, p, None, 1, 3, 65536, 65536, p, 1 << 20, None) == -4
    assert lib.fldr_splat_bwd(3, p, s4, p, s4, None, None, p, p, p, s4, p, None, None, 1, 3, 65536, 65536, None, 0, None) == -4
    assert lib.fldr_corr81_fwd(p, s4, p, s4, p, 70000, 4, 8, 8, None, 0, None) == -4
    assert lib.fldr_corr81_fwd(None, s4, p, s4, p, 1, 4, 8, 8, None, 0, None) == -1
    assert lib.fldr_corr81_fwd(p, s4, p, s4, p, 0, 4, 8, 8, None, 0, None) == -1
    assert lib.fldr_corr81_bwd(p, s4, p, s4, None, s4, p, p, 1, 4, 8, 8, None, 0, None) == -1
    assert lib.fldr_corr81_fwd_act(p, s4, p, s4, p, 10, 0.1, 1, 4, 8, 8, None) == -1      # sample stride < 81*H*W
    assert lib.fldr_corr81_fwd_act(p, s4, p, s4, p, 0, float("nan"), 1, 4, 8, 8, None) == -1
    assert lib.fldr_corr81_fwd_act(p, s4, None, s4, p, 0, 0.1, 1, 4, 8, 8, None) == -1
    # backward warp / splat metric (next row)
    assert lib.fldr_bwarp_fwd(None, s4, p, s4, p, 1, 3, 8, 8, 1, 0, None) == -1
    assert lib.fldr_bwarp_fwd(p, s4, p, s4, p, 1, 0, 8, 8, 1, 0, None) == -1
    assert lib.fldr_bwarp_fwd(p, s4, p, s4, p, 1, 3, 65536, 65536, 1, 0, None) == -4
    neg = (ctypes.c_int64 * 4)(64, 64, -8, 1)
    assert lib.fldr_bwarp_fwd(p, neg, p, s4, p, 1, 3, 8, 8, 1, 0, None) == -4       # flipped views are refused, not mis-read
    assert lib.fldr_bwarp_fwd(p, s4, p, s4, p, 1, 3, 8, 8, 1, 7, None) == -1          # unknown convention
    assert lib.fldr_bwarp_bwd(p, s4, p, s4, None, s4, p, p, 1, 3, 8, 8, 1, 0, None) == -1
    assert lib.fldr_bwarp_bwd(p, s4, p, s4, p, s4, p, p, 1, 3, 8, 8, 1, 2, None) == -1
    assert lib.fldr_warp_metric_fwd(p, s4, None, s4, p, s4, 1.0, p, 1, 3, 8, 8, 1, None) == -1
    assert lib.fldr_warp_metric_fwd(p, s4, p, s4, p, s4, 1.0, p, 1, 3, 8, 0, 1, None) == -1
    # occlusion softmax + blend (next row 2)
    six = (ctypes.c_void_p * 6)(*([16] * 6))
    s24 = (ctypes.c_int64 * 24)(*([64, 64, 8, 1] * 6))
    assert lib.fldr_occ_blend_fwd(p, s4, six, s24, p, 1, p, p, None, 1, 3, 8, 0, None) == -1
    assert lib.fldr_occ_blend_fwd(p, s4, six, s24, p, 1, None, p, None, 1, 3, 8, 8, None) == -1
    hole = (ctypes.c_void_p * 6)(16, 16, None, 16, 16, 16)
    assert lib.fldr_occ_blend_fwd(p, s4, hole, s24, p, 1, p, p, None, 1, 3, 8, 8, None) == -1
    assert lib.fldr_occ_blend_fwd(p, s4, six, s24, p, 1, p, p, None, 1, 3, 65536, 65536, None) == -4


def test_host_mirror_names_and_cpu_errors(lib):
    import fldr_vfi_b200.correlation as C
    import fldr_vfi_b200.softSplat as S
    for name in ("Softsplat", "FunctionSoftsplat", "_FunctionSoftsplat"):
        assert hasattr(S, name)
    for name in ("ModuleCorrelation", "FunctionCorrelation", "_FunctionCorrelation"):
        assert hasattr(C, name)
    assert S.Softsplat().strType == "softmax"                     # softSplat.py:356
    x, fl = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)
    with pytest.raises(NotImplementedError):                      # no CPU path, no fallback (softSplat.py:251-252)
        S.Softsplat()(x, fl)
    with pytest.raises(NotImplementedError):                      # correlation.py:343-344
        C.FunctionCorrelation(tensorFirst=x, tensorSecond=x)
    with pytest.raises(AssertionError):
        S.FunctionSoftsplat(x, fl, None, "nearest")
    import fldr_vfi_b200.warp as Wp
    with pytest.raises(NotImplementedError):                      # no CPU path for the warp row either
        Wp.bwarp(x, fl)
    with pytest.raises(NotImplementedError):
        Wp.splat_metric(x, x, fl, -1.9)
    import fldr_vfi_b200.blend as Bl
    with pytest.raises(NotImplementedError):
        Bl.occ_blend(torch.zeros(1, 6, 4, 4), torch.ones(1, dtype=torch.float64), torch.full((1, 1), 0.5), x, x, x, x, x, x)


def test_patch_bwarp_routes_only_what_the_kernel_covers(lib):
    """integrate.patch_bwarp swaps the method on a class and leaves CPU / autograd calls to the original."""
    import types
    from fldr_vfi_b200.integrate import patch_bwarp
    calls = []

    class DCTVFInet:
        def bwarp(self, x, flo, withmask=True, minus=False):
            calls.append("reference")
            return x
    mod = types.SimpleNamespace(DCTVFInet=DCTVFInet)
    original = patch_bwarp(mod)
    assert patch_bwarp(mod) is original                          # idempotent
    x, fl = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)
    assert DCTVFInet().bwarp(x, fl) is x and calls == ["reference"]      # CPU tensor -> the reference's own method
    DCTVFInet.bwarp = original
    assert not getattr(DCTVFInet.bwarp, "_fldr_b200_patched", False)


def test_keep_allocator_cache_is_reversible(lib):
    from fldr_vfi_b200.integrate import keep_allocator_cache
    original = keep_allocator_cache()
    try:
        assert keep_allocator_cache() is original and torch.cuda.empty_cache() is None
    finally:
        torch.cuda.empty_cache = original
    assert not getattr(torch.cuda.empty_cache, "_fldr_b200_patched", False)


def test_patch_pwc_backward_patches_instances(lib):
    from fldr_vfi_b200.integrate import patch_pwc_backward

    class Decoder(torch.nn.Module):
        def Backward(self, tensorInput, tensorFlow, g, p):
            return "reference"

    net = torch.nn.Sequential(Decoder(), torch.nn.Identity(), Decoder())
    assert patch_pwc_backward(net) == 2 and patch_pwc_backward(net) == 0          # idempotent
    x, fl = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)
    assert net[0].Backward(x, fl, {}, {}) == "reference"                          # CPU tensors -> the reference's method


def test_dropin_import_names_shadow_reference_modules(lib):
    """`from softSplat import Softsplat` (fLDRnet.py:22) and `from . import correlation` inside the OpticalFlow
    namespace package (PWCNet.py:4) resolve to the drop-ins when dropin/ is first on sys.path.  Importing the
    correlation drop-in must not touch CUDA (the reference does at correlation.py:7-8)."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, r'%s');"
        "import softSplat, OpticalFlow.correlation as oc;"
        "assert 'dropin' in softSplat.__file__ and 'dropin' in oc.__file__;"
        "assert softSplat.Softsplat.__module__ == 'fldr_vfi_b200.softSplat';"
        "assert oc.FunctionCorrelation.__module__ == 'fldr_vfi_b200.correlation';"
        "print('ok')" % os.path.join(ROOT, "fldr-vfi_b200", "dropin"))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr

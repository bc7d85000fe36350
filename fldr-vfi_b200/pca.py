"""Host mirror of the reference's block-PCA feature extraction (``pca_comp.py:473-528`` ``to_pca_diff``) over the C-ABI.

    to_pca_diff(im, params, args, mean, EV, mean_vec)      same signature and return value as the reference's function
    pca_features(im, mean, EV, mean_vec=None, out_dtype=torch.float64)

``im`` is ``[chan, H, W]`` float32 on the GPU (the two stacked frames of a batch, ``x_l[i].reshape(B*6, H, W)`` at
``fLDRnet.py:146``), ``mean`` / ``EV`` / ``mean_vec`` the float64 parameters of the model.  Returns ``[chan * ncomp, H/8, W/8]``
float64 like the reference (or float32 with ``out_dtype=torch.float32``: the ``.float()`` of the call site fused in).
Forward only: the shipped checkpoints train with ``noEVOptimization`` (fLDRnet.py:143-144); parameters that require grad raise.
"""
import ctypes

import torch

from . import _lib
from .softSplat import _cached_ws_bytes, _check_cuda_f32, _device_of, _stream_ptr, _workspace


def pca_features(im, mean, EV, mean_vec=None, out_dtype=torch.float64):
    if not im.is_cuda:
        raise NotImplementedError("pca_features: CPU tensors are not supported; there is no CPU fallback")
    _check_cuda_f32("im", im)
    for name, t in (("mean", mean), ("EV", EV), ("mean_vec", mean_vec)):
        if t is not None and t.dtype != torch.float64:
            raise TypeError(f"{name}: expected torch.float64 (the model keeps its PCA parameters in double), got {t.dtype}")
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (im, mean, EV, mean_vec)):
        raise NotImplementedError("pca_features is forward-only: call under torch.no_grad() (noEVOptimization)")
    if out_dtype not in (torch.float64, torch.float32):
        raise TypeError("out_dtype must be torch.float64 or torch.float32")
    chan, H, W = im.shape
    wiS = 8
    if H % wiS != 0 or W % wiS != 0:
        raise Exception("in to_pca_diff the image is not padded right." + str(H) + " " + str(W))       # pca_comp.py:486-487
    if im.stride(2) != 1:
        im = im.contiguous()
    ncomp = EV.shape[0]
    assert EV.shape[1] == wiS * wiS and mean.numel() == wiS * wiS and EV.stride(1) == 1
    mean = mean.contiguous()
    mv = None if mean_vec is None else mean_vec.contiguous()
    lib = _lib.lib()
    f32 = out_dtype == torch.float32
    out = torch.empty((chan * ncomp, H // wiS, W // wiS), dtype=out_dtype, device=im.device)
    ws_bytes = _cached_ws_bytes(lib.fldr_pca_features_workspace_bytes, chan, H, W, ncomp, int(f32))
    ws = _workspace(ws_bytes, im.device)
    strides = (ctypes.c_int64 * 3)(*im.stride())
    with _device_of(im):
        st = lib.fldr_pca_features_fwd(_lib.ptr(im), strides, _lib.ptr(mean), _lib.ptr(EV), EV.stride(0), _lib.ptr(mv), _lib.ptr(out),
                                       int(f32), chan, H, W, ncomp, _lib.ptr(ws), ws_bytes, _stream_ptr(im.device))
    _lib.check(st)
    return out


def to_pca_diff(im, params, args, mean, EV, mean_vec):
    """Drop-in for ``pca_comp.to_pca_diff`` (same arguments; ``params.wiS`` must be 8 and ``components_fraction`` must select
    the ``EV.shape[0]`` rows passed in, as at fLDRnet.py:146)."""
    assert params.wiS == 8, "the block size of the shipped models"
    assert int(params.wiS * params.wiS * params.components_fraction) == EV.shape[0]
    return pca_features(torch.as_tensor(im), mean, EV, mean_vec if args.mean_vector_norm else None)

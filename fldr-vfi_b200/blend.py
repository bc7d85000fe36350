"""Host mirror of fLDRnet's occlusion softmax + image synthesis over the sm_100a C-ABI library.

  occ_blend(refine_out, T_param, t_value, warped_img0, warped_img1, im0_tot, im1_tot, x0, x1, return_occ0=False)

computes what /root/reference/fLDRnet.py:510-524 computes - ``out_l`` (float64, like the reference: ``T_param`` is a
float64 Parameter and promotes the whole expression) and optionally ``occ_0_l`` (line 512) - in one kernel launch,
with no host synchronisation (``T_param`` and ``t_value`` are read on the device).  Differentiable: the forward is the
fused kernel; the backward re-evaluates the reference's expression (fLDRnet.py:510-524, restated in ``_blend_expression``) with
torch's autograd on the device, so every gradient (logits, the six images, ``T_param`` when ``TOptimization`` is on, ``t_value``)
is the one autograd derives for the reference itself.  ``return_occ0`` is forward-only.
"""
import ctypes

import torch

from . import _lib
from .softSplat import _check_cuda_f32, _device_of, _stream_ptr


def _blend_expression(refine_out, T_param, t_value, images):
    """fLDRnet.py:510-524 as torch operators (used by the backward only).  float64 from the division by T_param on."""
    occ = torch.softmax(refine_out[:, 0:6] / T_param, dim=1)                                             # 511
    t = t_value.view(-1, 1, 1, 1)
    a = [((1 - t) if k % 2 == 0 else t) * occ[:, k:k + 1] for k in range(6)]
    divisor = a[0] + a[1] + a[2] + a[3]                                                                  # 517
    out = a[0] * images[0] + a[1] * images[1]                                                            # 518
    out = out + (a[2] * images[2] + a[3] * images[3])                                                    # 520
    out = out + (a[4] * images[4] + a[5] * images[5])                                                    # 521
    divisor = divisor + (a[4] + a[5])                                                                    # 522
    return out / divisor                                                                                 # 524


class _FunctionOccBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, refine_out, T_param, t_value, *images):
        ctx.save_for_backward(refine_out, T_param, t_value, *images)
        return _occ_blend_forward(refine_out, T_param, t_value, images, False)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        leaves = [t.detach().requires_grad_(need) for t, need in zip(saved, ctx.needs_input_grad)]
        with torch.enable_grad():
            out = _blend_expression(leaves[0], leaves[1], leaves[2], leaves[3:])
        wrt = [t for t in leaves if t.requires_grad]
        grads = iter(torch.autograd.grad(out, wrt, grad_out))
        return tuple(next(grads) if t.requires_grad else None for t in leaves)


def occ_blend(refine_out, T_param, t_value, warped_img0, warped_img1, im0_tot, im1_tot, x0, x1, return_occ0=False):
    images = (warped_img0, warped_img1, im0_tot, im1_tot, x0, x1)            # order of fLDRnet.py:518-521
    if torch.is_grad_enabled() and any(t.requires_grad for t in (refine_out, T_param, t_value) + images):
        if return_occ0:
            raise NotImplementedError("occ_blend(return_occ0=True) is forward-only: take occ_0 from softmax(refine_out / T_param)")
        _check_inputs(refine_out, T_param, t_value, images)
        return _FunctionOccBlend.apply(refine_out, T_param, t_value, *images)
    return _occ_blend_forward(refine_out, T_param, t_value, images, return_occ0)


def _check_inputs(refine_out, T_param, t_value, images):
    if not refine_out.is_cuda:
        raise NotImplementedError()
    _check_cuda_f32("refine_out", refine_out)
    for i, im in enumerate(images):
        _check_cuda_f32(f"image {i}", im)
    _check_cuda_f32("t_value", t_value)
    if not (T_param.is_cuda and T_param.dtype == torch.float64 and T_param.numel() == 1):
        raise TypeError("T_param must be a CUDA float64 tensor with one element (fLDRnet.py:357)")


def _occ_blend_forward(refine_out, T_param, t_value, images, return_occ0):
    _check_inputs(refine_out, T_param, t_value, images)
    N, C, H, W = images[0].shape
    assert refine_out.shape[0] == N and refine_out.shape[1] >= 6 and refine_out.shape[2:] == (H, W)
    assert all(im.shape == (N, C, H, W) for im in images)
    t_flat = t_value.reshape(-1)
    assert t_flat.numel() == N, "t_value holds one value per sample (fLDRnet.py:112-119)"
    lib = _lib.lib()
    out = torch.empty((N, C, H, W), dtype=torch.float64, device=refine_out.device)
    occ0 = torch.empty((N, 1, H, W), dtype=torch.float64, device=refine_out.device) if return_occ0 else None
    ptrs = (ctypes.c_void_p * 6)(*[im.data_ptr() for im in images])
    strides = (ctypes.c_int64 * 24)(*[s for im in images for s in im.stride()])
    with _device_of(refine_out):
        st = lib.fldr_occ_blend_fwd(_lib.ptr(refine_out), _lib.strides(refine_out), ptrs, strides, _lib.ptr(t_flat),
                                    t_flat.stride(0) if N > 1 else 0, ctypes.c_void_p(T_param.data_ptr()),
                                    ctypes.c_void_p(out.data_ptr()), None if occ0 is None else ctypes.c_void_p(occ0.data_ptr()),
                                    N, C, H, W, _stream_ptr(refine_out.device))
    _lib.check(st)
    return (out, occ0) if return_occ0 else out

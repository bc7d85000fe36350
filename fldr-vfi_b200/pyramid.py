"""Host mirror of the reference's input pyramid (``main.py:855-856`` test, ``562-563`` train) over the C-ABI.

The reference builds level ``i > 0`` of ``input_gpu`` with ``F.interpolate(..., scale_factor=scales[0] / scales[i],
mode='bicubic', align_corners=args.align_cornerse)`` on the CPU from the full-resolution padded frames and copies every level
to the device.  ``input_pyramid`` takes the frames already on the GPU and writes all levels with one call (one launch for
the shipped scale presets):

    input_gpu = input_pyramid(input_frames.to(device), args.scales, args.S_tst, align_corners=args.align_cornerse)

``input_frames`` is ``[B, C, T, H, W]`` float32; the result is the list the reference builds (level 0 is the input itself,
level ``i`` is ``[B, C, T, floor(H f_i), floor(W f_i)]``).  Forward only, like the reference (data preparation, no graph).
"""
import ctypes
import math

import torch

from . import _lib
from .softSplat import _check_cuda_f32, _device_of, _stream_ptr


def bicubic_levels(planes, scale_factors, align_corners=False):
    """``planes`` ``[..., H, W]`` float32 CUDA (unit pixel stride, uniform plane stride) -> one contiguous ``[..., h_i, w_i]``
    tensor per scale factor."""
    if not planes.is_cuda:
        raise NotImplementedError("bicubic_levels: CPU tensors are not supported; there is no CPU fallback")
    _check_cuda_f32("planes", planes)
    if planes.dim() < 2:
        raise ValueError("planes: expected [..., H, W]")
    H, W = planes.shape[-2:]
    x = planes.reshape(-1, H, W)
    if x.stride(2) != 1:
        x = x.contiguous()
    n = len(scale_factors)
    sizes = [(int(math.floor(float(H) * f)), int(math.floor(float(W) * f))) for f in scale_factors]     # F.interpolate's size rule
    if any(h < 1 or w < 1 for h, w in sizes):
        raise RuntimeError("Input and output sizes should be greater than 0")                           # torch's own message
    outs = [torch.empty(tuple(planes.shape[:-2]) + s, dtype=torch.float32, device=planes.device) for s in sizes]
    if n == 0 or x.shape[0] == 0:
        return outs
    factors = (ctypes.c_double * n)(*[float(f) for f in scale_factors])
    ptrs = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    with _device_of(x):
        st = _lib.lib().fldr_bicubic_pyramid_fwd(_lib.ptr(x), x.stride(0), x.stride(1), x.shape[0], H, W, n, factors,
                                                 int(bool(align_corners)), ptrs, _stream_ptr(x.device))
    _lib.check(st)
    return outs


def input_pyramid(input_frames, scales, n_levels, align_corners=False):
    """main.py:855-856 with the frames on the device: ``[input_frames] + [bicubic level i for i in 1..n_levels]``."""
    if input_frames.dim() != 5:
        raise ValueError("input_frames: expected [B, C, T, H, W]")
    factors = [scales[0] / scales[i] for i in range(1, n_levels + 1)]
    return [input_frames] + bicubic_levels(input_frames, factors, align_corners)

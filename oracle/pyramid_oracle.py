"""NumPy restatement of the input pyramid of the reference (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates ``/root/reference/main.py:855-856`` (test) / ``562-563`` (train): level ``i > 0`` of ``input_gpu`` is
``F.interpolate(frames, scale_factor=scales[0] / scales[i], mode='bicubic', align_corners=args.align_cornerse)`` of the
FULL-resolution padded frames, computed on the CPU and then copied to the device.  The interpolation itself lives in a
third-party dependency that is not part of the checkout - PyTorch (the image has 2.11; the reference pins none) - so its
published CPU algorithm is restated here operation by operation in float32:

  * ``aten/src/ATen/native/UpSample.h``: ``area_pixel_compute_scale`` (``1 / scale_factor`` when a scale factor is given,
    ``(in-1)/(out-1)`` with align_corners), ``area_pixel_compute_source_index`` (``scale * (dst + 0.5) - 0.5``, not clamped
    for cubic), ``guard_index_and_lambda``, ``get_cubic_upsample_coefficients`` with ``A = -0.75``;
  * ``aten/src/ATen/native/cpu/UpSampleKernel.cpp`` ``HelperInterpCubic`` + ``Interpolate<2>``: four taps per axis at
    ``clamp(floor(src) - 1 + j, 0, size - 1)``, the horizontal axis innermost, partial sums accumulated left to right.

Pinned against ``tests/golden/pyramid_*.npz``, which ``tests/golden/make_golden.py --pyramid-only`` produced by executing the
reference's own list comprehension (main.py:855-856, lifted by ``ast``) on the CPU.
"""
import math

import numpy as np

F32 = np.float32
A = F32(-0.75)


def _conv1(x):            # UpSample.h cubic_convolution1:  ((A + 2) x - (A + 3)) x x + 1
    return ((A + F32(2)) * x - (A + F32(3))) * x * x + F32(1)


def _conv2(x):            # UpSample.h cubic_convolution2:  ((A x - 5A) x + 8A) x - 4A
    return ((A * x - F32(5) * A) * x + F32(8) * A) * x - F32(4) * A


def axis_taps(in_size, out_size, scale_factor, align_corners):
    """Tap indices ``[out, 4]`` (int64, clamped) and weights ``[out, 4]`` (float32) of one axis."""
    if align_corners:
        scale = F32(in_size - 1) / F32(out_size - 1) if out_size > 1 else F32(0)
    else:
        scale = F32(1.0 / scale_factor) if scale_factor is not None and scale_factor > 0 else F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = scale * dst if align_corners else scale * (dst + F32(0.5)) - F32(0.5)
    fl = np.floor(src)
    idx = np.minimum(fl.astype(np.int64), in_size - 1)                      # guard_index_and_lambda
    lam = np.minimum(np.maximum(src - idx.astype(F32), F32(0)), F32(1)).astype(F32)
    one_m = (F32(1) - lam).astype(F32)
    w = np.stack([_conv2(lam + F32(1)), _conv1(lam), _conv1(one_m), _conv2(one_m + F32(1))], 1).astype(F32)
    taps = np.clip(idx[:, None] + np.arange(-1, 3)[None, :], 0, in_size - 1)
    return taps, w


def output_size(in_size, scale_factor):
    return int(math.floor(float(in_size) * scale_factor))                   # torch.nn.functional.interpolate


def bicubic_resize(x, scale_factor, align_corners=False):
    """``x`` ``[..., H, W]`` float32 -> ``[..., floor(H s), floor(W s)]`` float32."""
    x = np.asarray(x, dtype=F32)
    H, W = x.shape[-2:]
    oh, ow = output_size(H, scale_factor), output_size(W, scale_factor)
    ty, wy = axis_taps(H, oh, scale_factor, align_corners)
    tx, wx = axis_taps(W, ow, scale_factor, align_corners)
    rows = x[..., :, tx]                                                    # [..., H, ow, 4]
    h = rows[..., 0] * wx[:, 0]
    for j in range(1, 4):
        h = (h + rows[..., j] * wx[:, j]).astype(F32)
    v = h[..., ty, :]                                                       # [..., oh, 4, ow]
    out = v[..., 0, :] * wy[:, 0, None]
    for j in range(1, 4):
        out = (out + v[..., j, :] * wy[:, j, None]).astype(F32)
    return out.astype(F32)


def input_pyramid(input_frames, scales, n_levels, align_corners=False):
    """main.py:855-856: ``input_frames`` ``[B, C, T, H, W]``; returns ``n_levels + 1`` arrays ``[B, C, T, h_i, w_i]``, level 0
    being the frames themselves (the permute / reshape pair of the reference only routes planes: interpolation is per
    plane)."""
    out = [np.asarray(input_frames, dtype=F32)]
    for i in range(1, n_levels + 1):
        out.append(bicubic_resize(input_frames, scales[0] / scales[i], align_corners))
    return out

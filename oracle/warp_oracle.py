"""CPU restatement of fLDRnet's backward warp and splat metric.  TEST INFRASTRUCTURE ONLY:
nothing in the product path may import this module (tests/, __graft_entry__.smoke and bench.py's cpu_baseline leg only).

Follows, explicitly and without calling ``grid_sample``:

* ``DCTVFInet.bwarp(x, flo, withmask)``                      fLDRnet.py:546-581
    grid = (x + u, y + v)                                     :556-562
    gx = 2*X/max(W-1,1) - 1,  gy = 2*Y/max(H-1,1) - 1         :565-566   (align_corners=True style normalisation ...)
    output = grid_sample(x, grid)                             :568       (... fed to the align_corners=False default:
                                                                          ix = ((gx+1)*W - 1)/2 = X*W/(W-1) - 0.5, a quirk
                                                                          we preserve; one rounding, see _unnormalise)
    mask = grid_sample(ones, grid); mask<0.999 -> 0; mask>0 -> 1   :569-574
    return output*mask if withmask else output                :578-581
  bilinear, zero padding: taps (floor(ix), floor(iy)) + {0,1}^2 with weights (1-tx)(1-ty) ..., taps outside the
  frame contribute nothing (torch grid_sampler, mode='bilinear', padding_mode='zeros').
* the splat metric  z = mean_c( z_alpha * |x_ref - bwarp(x_src, flo)| )   fLDRnet.py:442-446

Pinned against tests/golden/warp_*.npz, which tests/golden/make_golden.py produced by executing the reference's own
``bwarp`` source (extracted from /root/reference/fLDRnet.py, run on the CPU) and the expression of line 443.
"""
import torch


def _unnormalise(g, size):
    """grid_sample's align_corners=False un-normalisation ((g + 1) * size - 1) / 2.  torch evaluates it with ONE rounding
    (a fused multiply-add, on the CPU's vector path and on CUDA alike - established against the golden vectors, which
    the unfused form misses by an ulp of the coordinate); emulated exactly here by evaluating in float64 (the 24 x 13
    bit product is exact there) and rounding once."""
    return (((g + 1.0).double() * size - 1.0) / 2.0).float()


def _div_scalar(t, b, cuda_semantics):
    """``tensor / python_scalar`` as torch evaluates it: a true division on the CPU, a multiplication by the float32
    reciprocal ``1/b`` on CUDA (BinaryDivTrueKernel.cu) - one ulp apart, and the reference runs on CUDA."""
    if cuda_semantics:
        return t * (torch.tensor(1.0) / torch.tensor(float(b)))
    return t / b


def _source_index(coord, size, cuda_semantics=False):
    """fLDRnet.py:565-566 in float32, operation by operation, followed by grid_sample's un-normalisation."""
    g = _div_scalar(2.0 * coord, max(size - 1, 1), cuda_semantics) - 1.0
    return _unnormalise(g, size)


def _linspace_pm1(n):
    """torch.linspace(-1, 1, n) on the CPU, restated: float32 step, one fused multiply-add per element, mirrored halves
    (the fused rounding is emulated exactly by evaluating in float64 and rounding once: step*i has at most 48 bits)."""
    if n == 1:
        return torch.tensor([-1.0])
    step = (torch.tensor(2.0) / torch.tensor(float(n - 1))).double()
    i = torch.arange(n, dtype=torch.float64)
    lo = (-1.0 + step * i).float()
    hi = (1.0 - step * (n - 1 - i)).float()
    return torch.where(torch.arange(n) < n // 2, lo, hi)


def pwc_backward(x, flo, return_mask=False, cuda_semantics=False):
    """PWC-Net's Backward (OpticalFlow/PWCNet.py:116-143): grid = linspace(-1,1) per axis (:117-130), flow divided by
    (size-1)/2 (:134-135), grid_sample bilinear / zeros / align_corners=False default of the input with a ones channel
    (:136-138), mask = [ones channel > 0.999] (:140-141), output * mask (:143)."""
    x = x.float()
    flo = flo.float()
    B, C, H, W = x.shape
    gx = _linspace_pm1(W).view(1, 1, W) + _div_scalar(flo[:, 0], (W - 1.0) / 2.0, cuda_semantics)
    gy = _linspace_pm1(H).view(1, H, 1) + _div_scalar(flo[:, 1], (H - 1.0) / 2.0, cuda_semantics)
    return _sample(x, _unnormalise(gx, W), _unnormalise(gy, H), True, return_mask, strict=True)


def bwarp(x, flo, withmask=True, return_mask=False, cuda_semantics=False):
    """``cuda_semantics``: evaluate ``/ scalar`` the way torch's CUDA kernels do (see _div_scalar).  False reproduces the
    CPU run that produced tests/golden/warp_*.npz; True is what the reference computes on a GPU and what the sm_100a
    kernel follows."""
    x = x.float()
    flo = flo.float()
    B, C, H, W = x.shape
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(B, H, W)
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    ix = _source_index(xx + flo[:, 0], W, cuda_semantics)
    iy = _source_index(yy + flo[:, 1], H, cuda_semantics)
    return _sample(x, ix, iy, withmask, return_mask, strict=False)


def _sample(x, ix, iy, withmask, return_mask, strict):
    B, C, H, W = x.shape
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    taps = [(x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
            (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))]
    out = torch.zeros_like(x)
    msum = torch.zeros(B, H, W)
    flat = x.reshape(B, C, H * W)
    for tx, ty, w in taps:
        ok = (tx >= 0) & (tx <= W - 1) & (ty >= 0) & (ty <= H - 1)          # False for NaN / inf coordinates too
        zero = torch.zeros_like(tx)
        idx = (torch.where(ok, ty, zero) * W + torch.where(ok, tx, zero)).long().view(B, 1, H * W).expand(B, C, H * W)
        v = torch.gather(flat, 2, idx).view(B, C, H, W)
        wk = torch.where(ok, w, torch.zeros_like(w))
        out = out + v * wk.unsqueeze(1)
        msum = msum + wk
    # fLDRnet: <0.999 -> 0, every survivor (>0) -> 1;  PWC-Net: >0.999 -> 1, everything else (<1) -> 0
    mask = ((msum > 0.999) if strict else (msum >= 0.999)).float().unsqueeze(1)
    res = out * mask if withmask else out
    return (res, msum) if return_mask else res


def warp_metric(x_ref, x_src, flo, alpha, withmask=True, cuda_semantics=False):
    """z = mean_c(alpha * |x_ref - bwarp(x_src, flo)|), keepdim (fLDRnet.py:442-443).  torch.mean divides the sum by C on
    the CPU and multiplies it by the float32 factor 1/C on CUDA (ReduceMomentKernel.cu)."""
    w = bwarp(x_src, flo, withmask, cuda_semantics=cuda_semantics)
    t = float(alpha) * torch.abs(x_ref.float() - w)
    if not cuda_semantics:
        return torch.mean(t, dim=1, keepdim=True)
    C = t.shape[1]
    s = t[:, 0:1]
    for c in range(1, C):
        s = s + t[:, c:c + 1]
    return s * (torch.tensor(1.0) / torch.tensor(float(C)))

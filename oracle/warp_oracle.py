"""CPU restatement of fLDRnet's backward warp and splat metric.  TEST INFRASTRUCTURE ONLY:
nothing in the product path may import this module (tests/, __graft_entry__.smoke and bench.py's cpu_baseline leg only).

Follows, explicitly and without calling ``grid_sample``:

* ``DCTVFInet.bwarp(x, flo, withmask)``                      fLDRnet.py:546-581
    grid = (x + u, y + v)                                     :556-562
    gx = 2*X/max(W-1,1) - 1,  gy = 2*Y/max(H-1,1) - 1         :565-566   (align_corners=True style normalisation ...)
    output = grid_sample(x, grid)                             :568       (... fed to the align_corners=False default:
                                                                          ix = ((gx+1)*W - 1)/2 = X*W/(W-1) - 0.5, a quirk
                                                                          we preserve)
    mask = grid_sample(ones, grid); mask<0.999 -> 0; mask>0 -> 1   :569-574
    return output*mask if withmask else output                :578-581
  bilinear, zero padding: taps (floor(ix), floor(iy)) + {0,1}^2 with weights (1-tx)(1-ty) ..., taps outside the
  frame contribute nothing (torch grid_sampler, mode='bilinear', padding_mode='zeros').
* the splat metric  z = mean_c( z_alpha * |x_ref - bwarp(x_src, flo)| )   fLDRnet.py:442-446

Pinned against tests/golden/warp_*.npz, which tests/golden/make_golden.py produced by executing the reference's own
``bwarp`` source (extracted from /root/reference/fLDRnet.py, run on the CPU) and the expression of line 443.
"""
import torch


def _source_index(coord, size):
    """fLDRnet.py:565-566 followed by grid_sample's align_corners=False un-normalisation, in fp32 and in that order."""
    g = 2.0 * coord / max(size - 1, 1) - 1.0
    return ((g + 1.0) * size - 1.0) / 2.0


def bwarp(x, flo, withmask=True, return_mask=False):
    x = x.float()
    flo = flo.float()
    B, C, H, W = x.shape
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(B, H, W)
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    ix = _source_index(xx + flo[:, 0], W)
    iy = _source_index(yy + flo[:, 1], H)
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
    mask = (msum >= 0.999).float().unsqueeze(1)          # <0.999 -> 0, every survivor (>0) -> 1
    res = out * mask if withmask else out
    return (res, msum) if return_mask else res


def warp_metric(x_ref, x_src, flo, alpha, withmask=True):
    """z = mean_c(alpha * |x_ref - bwarp(x_src, flo)|), keepdim (fLDRnet.py:442-443)."""
    w = bwarp(x_src, flo, withmask)
    return torch.mean(float(alpha) * torch.abs(x_ref.float() - w), dim=1, keepdim=True)

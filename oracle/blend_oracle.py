"""CPU restatement of fLDRnet's occlusion softmax + six-way image synthesis.
TEST INFRASTRUCTURE ONLY: nothing in the product path may import this module.

Follows fLDRnet.py:510-524 term by term, with the softmax written out (no ``F.softmax``):

    occ = softmax(refine_out[:, 0:6] / T_param, dim=1)          511   T_param is a float64 Parameter of shape [1]
                                                                      (357), so everything from here on is float64
    a0 = (1-t)*occ0  a1 = t*occ1  a2 = (1-t)*occ2  a3 = t*occ3  a4 = (1-t)*occ4  a5 = t*occ5
                                                                      ((1-t) is computed in float32, t_value's dtype)
    divisor = ((a0 + a1) + a2) + a3                              517
    out  = a0*warped0 + a1*warped1                               518
    out += a2*im0_tot + a3*im1_tot                               520
    out += a4*x0 + a5*x1                                         521
    divisor += a4 + a5                                           522
    out /= divisor                                               524
    occ_0 = occ[:, 0:1]                                          512

Pinned against tests/golden/blend_*.npz, produced by tests/golden/make_golden.py by executing those very source lines of
/root/reference/fLDRnet.py on the CPU.
"""
import torch


def occ_blend(refine_out, T_param, t_value, warped0, warped1, im0_tot, im1_tot, x0, x1):
    """All images [N,C,H,W] float32, refine_out [N,>=6,H,W] float32, T_param float64 [1], t_value float32 [N,1,1,1] (or
    [N,1]).  Returns (out float64 [N,C,H,W], occ_0 float64 [N,1,H,W])."""
    e = refine_out[:, 0:6].double() / T_param.double().reshape(())
    m = e.max(dim=1, keepdim=True).values
    p = torch.exp(e - m)
    s = p[:, 0:1]
    for k in range(1, 6):
        s = s + p[:, k:k + 1]
    occ = p / s
    t32 = t_value.float().view(-1, 1, 1, 1)
    one_minus_t = (1 - t32).double()                     # float32 subtraction, then promoted
    t = t32.double()
    a = [(one_minus_t if k % 2 == 0 else t) * occ[:, k:k + 1] for k in range(6)]
    divisor = ((a[0] + a[1]) + a[2]) + a[3]
    out = a[0] * warped0 + a[1] * warped1
    out = out + (a[2] * im0_tot + a[3] * im1_tot)
    out = out + (a[4] * x0 + a[5] * x1)
    divisor = divisor + (a[4] + a[5])
    return out / divisor, occ[:, 0:1]

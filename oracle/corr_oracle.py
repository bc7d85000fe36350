"""Torch-CPU restatement of the reference correlation op (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates ``/root/reference/OpticalFlow/correlation.py``: 81-channel cost volume,
displacements dy,dx in [-4,4], zero padding 4, mean over channels, output channel
``(dy+4)*9+(dx+4)`` (correlation.py:81-82,106-108).
"""
import torch
import torch.nn.functional as F

PAD = 4
D = 9


def rearrange(x):
    """``kernel_Correlation_rearrange`` (correlation.py:17-42): NCHW -> zero-padded NHWC."""
    return F.pad(x, (PAD, PAD, PAD, PAD)).permute(0, 2, 3, 1).contiguous()


def correlation_fwd(first, second):
    """``kernel_Correlation_updateOutput`` (correlation.py:44-112).

    ``out[b,(dy+4)*9+(dx+4),y,x] = (1/C) * sum_c f1[b,c,y,x] * f2z[b,c,y+dy,x+dx]``.
    """
    B, C, H, W = first.shape
    assert second.shape == first.shape
    f2p = F.pad(second, (PAD, PAD, PAD, PAD))
    out = torch.empty(B, D * D, H, W, dtype=first.dtype)
    for p in range(-PAD, PAD + 1):          # s2p, vertical (correlation.py:82)
        for o in range(-PAD, PAD + 1):      # s2o, horizontal (correlation.py:81)
            sh = f2p[:, :, PAD + p:PAD + p + H, PAD + o:PAD + o + W]
            out[:, (p + PAD) * D + (o + PAD)] = (first * sh).sum(1) / C
    return out


def correlation_grad_first(second, grad_out):
    """``kernel_Correlation_updateGradFirst`` (correlation.py:114-176).

    ``gF1[b,c,y,x] = (1/C) * sum_{dy,dx} gOut[b,op,y,x] * f2z[b,c,y+dy,x+dx]``.
    """
    B, C, H, W = second.shape
    f2p = F.pad(second, (PAD, PAD, PAD, PAD))
    g = torch.zeros(B, C, H, W, dtype=second.dtype)
    for p in range(-PAD, PAD + 1):
        for o in range(-PAD, PAD + 1):
            sh = f2p[:, :, PAD + p:PAD + p + H, PAD + o:PAD + o + W]
            g += grad_out[:, (p + PAD) * D + (o + PAD)].unsqueeze(1) * sh
    return g / C


def correlation_grad_second(first, grad_out):
    """``kernel_Correlation_updateGradSecond`` (correlation.py:178-242).

    ``gF2[b,c,y,x] = (1/C) * sum_{dy,dx} [0<=y-dy<H][0<=x-dx<W] gOut[b,op,y-dy,x-dx] * f1[b,c,y-dy,x-dx]``.
    """
    B, C, H, W = first.shape
    f1p = F.pad(first, (PAD, PAD, PAD, PAD))
    gp = F.pad(grad_out, (PAD, PAD, PAD, PAD))
    g = torch.zeros(B, C, H, W, dtype=first.dtype)
    for p in range(-PAD, PAD + 1):
        for o in range(-PAD, PAD + 1):
            f1s = f1p[:, :, PAD - p:PAD - p + H, PAD - o:PAD - o + W]
            gos = gp[:, (p + PAD) * D + (o + PAD), PAD - p:PAD - p + H, PAD - o:PAD - o + W]
            g += gos.unsqueeze(1) * f1s
    return g / C


class _Correlation(torch.autograd.Function):
    """Autograd pairing identical to ``_FunctionCorrelation`` (correlation.py:294-409)."""

    @staticmethod
    def forward(ctx, first, second):
        ctx.save_for_backward(first, second)
        return correlation_fwd(first, second)

    @staticmethod
    def backward(ctx, grad_out):
        first, second = ctx.saved_tensors
        g1 = correlation_grad_first(second, grad_out) if ctx.needs_input_grad[0] else None
        g2 = correlation_grad_second(first, grad_out) if ctx.needs_input_grad[1] else None
        return g1, g2


def function_correlation(tensorFirst, tensorSecond):
    """``FunctionCorrelation`` (correlation.py:415-416)."""
    return _Correlation.apply(tensorFirst, tensorSecond)

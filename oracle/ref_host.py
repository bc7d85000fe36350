"""ctypes driver for ``oracle/_ref/libref_host.so`` - the reference's own kernel text run on
host cores (TEST INFRASTRUCTURE - see oracle/__init__.py and oracle/build_ref.py).

The torch-level glue around the kernels (allocation, the per-mode pre/post-processing) is
the restatement in ``splat_oracle.function_softsplat`` with the raw splat swapped for the
reference kernel, so a value computed here went through the reference's arithmetic for
everything the reference implements as a kernel.
"""
import ctypes
import os

import numpy as np
import torch

from . import build_ref

_LIB = None
_F = ctypes.POINTER(ctypes.c_float)


def available():
    return os.path.exists(os.path.join(build_ref.OUT, "libref_host.so")) or os.path.isdir(build_ref.REF)


def lib():
    global _LIB
    if _LIB is None:
        so = build_ref.build(verbose=False)
        if so is None:
            raise RuntimeError("oracle/_ref/libref_host.so missing and /root/reference not present")
        _LIB = ctypes.CDLL(so)
        _LIB.ref_num_threads.restype = ctypes.c_int
    return _LIB


def _p(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), _F)


def num_threads():
    return int(lib().ref_num_threads())


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))


def splat_update_output(inp, flow):
    inp = inp.contiguous(); flow = flow.contiguous()
    N, C, H, W = inp.shape
    out = torch.empty_like(inp)
    lib().ref_splat_update_output(_p(inp), _p(flow), _p(out), N, C, H, W)
    return out


def splat_update_grad_input(inp, flow, gout):
    inp = inp.contiguous(); flow = flow.contiguous(); gout = gout.contiguous()
    N, C, H, W = inp.shape
    gin = torch.zeros_like(inp)
    lib().ref_splat_update_grad_input(_p(inp), _p(flow), _p(gout), _p(gin), N, C, H, W)
    return gin


def splat_update_grad_flow(inp, flow, gout):
    inp = inp.contiguous(); flow = flow.contiguous(); gout = gout.contiguous()
    N, C, H, W = inp.shape
    gfl = torch.zeros_like(flow)
    lib().ref_splat_update_grad_flow(_p(inp), _p(flow), _p(gout), _p(gfl), N, C, H, W)
    return gfl


class RefSplatRaw(torch.autograd.Function):
    """``_FunctionSoftsplat`` with the reference kernels on the host."""

    @staticmethod
    def forward(ctx, inp, flow):
        ctx.save_for_backward(inp, flow)
        return splat_update_output(inp, flow)

    @staticmethod
    def backward(ctx, gout):
        inp, flow = ctx.saved_tensors
        gi = splat_update_grad_input(inp, flow, gout) if ctx.needs_input_grad[0] else None
        gf = splat_update_grad_flow(inp, flow, gout) if ctx.needs_input_grad[1] else None
        return gi, gf


def function_softsplat(tenInput, tenFlow, tenMetric, strType):
    from . import splat_oracle
    return splat_oracle.function_softsplat(tenInput, tenFlow, tenMetric, strType, raw=RefSplatRaw.apply)


def corr_rearrange(x):
    x = x.contiguous()
    B, C, H, W = x.shape
    rb = torch.empty(B, H + 8, W + 8, C, dtype=torch.float32)
    lib().ref_corr_rearrange(_p(x), _p(rb), B, C, H, W)
    return rb


def corr_update_output(first, second):
    B, C, H, W = first.shape
    rb0, rb1 = corr_rearrange(first), corr_rearrange(second)
    out = torch.empty(B, 81, H, W, dtype=torch.float32)
    lib().ref_corr_update_output(_p(rb0), _p(rb1), _p(out), B, C, H, W)
    return out


def corr_update_grads(first, second, gout):
    B, C, H, W = first.shape
    rb0, rb1 = corr_rearrange(first), corr_rearrange(second)
    gout = gout.contiguous()
    g1 = torch.zeros_like(first)
    g2 = torch.zeros_like(first)
    lib().ref_corr_update_grad_first(_p(rb0), _p(rb1), _p(gout), _p(g1), B, C, H, W)
    lib().ref_corr_update_grad_second(_p(rb0), _p(rb1), _p(gout), _p(g2), B, C, H, W)
    return g1, g2

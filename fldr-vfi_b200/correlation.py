"""Host mirror of the reference's ``OpticalFlow/correlation.py`` over the sm_100a C-ABI library.

Same names and call shapes (file:line = /root/reference/OpticalFlow/correlation.py):

  FunctionCorrelation(tensorFirst, tensorSecond)         415-416   (keyword-callable, PWCNet.py:188,198)
  ModuleCorrelation().forward(tensorFirst, tensorSecond) 421-428
  _FunctionCorrelation                                   294-409

Deliberate divergences: importing this module does not touch CUDA (the reference captures
``torch.cuda.current_stream()`` at import, 7-8, and is pinned to that stream forever); launches go to the
current stream at call time; no ``rbot0/rbot1`` NHWC scratch copies are made or saved (297-300); the backward
is one batched launch per gradient instead of one per sample (365, 385).
"""
import torch

from . import _lib
from .softSplat import _cached_ws_bytes, _check_cuda_f32, _device_of, _stream_ptr, _workspace


def _corr_forward(first, second):
    if not first.is_cuda:
        raise NotImplementedError()                      # correlation.py:343-344
    _check_cuda_f32("first", first)
    _check_cuda_f32("second", second)
    assert (first.is_contiguous() == True)               # correlation.py:302
    assert (second.is_contiguous() == True)              # correlation.py:303
    assert first.shape == second.shape
    ext = _lib.ext()
    if ext is not None:
        return ext.corr81_fwd(first, second)
    lib = _lib.lib()
    B, C, H, W = first.shape
    output = torch.empty((B, 81, H, W), dtype=torch.float32, device=first.device)
    ws_bytes = _cached_ws_bytes(lib.fldr_corr81_fwd_workspace_bytes, B, C, H, W)
    ws = _workspace(ws_bytes, first.device)
    with _device_of(first):
        st = lib.fldr_corr81_fwd(_lib.ptr(first), _lib.strides(first), _lib.ptr(second), _lib.strides(second),
                                 _lib.ptr(output), B, C, H, W, _lib.ptr(ws), ws_bytes, _stream_ptr(first.device))
    _lib.check(st)
    return output


class _FunctionCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(self, first, second):
        output = _corr_forward(first, second)
        self.save_for_backward(first, second)
        return output

    @staticmethod
    def backward(self, gradOutput):
        first, second = self.saved_tensors
        _check_cuda_f32("gradOutput", gradOutput)
        gradOutput = gradOutput.contiguous()                 # reference asserts contiguity (356); we accept any layout
        ext = _lib.ext()
        if ext is not None:
            g1, g2 = ext.corr81_bwd(first, second, gradOutput, bool(self.needs_input_grad[0]), bool(self.needs_input_grad[1]))
            return g1, g2
        lib = _lib.lib()
        B, C, H, W = first.shape
        gradFirst = torch.empty_like(first) if self.needs_input_grad[0] else None
        gradSecond = torch.empty_like(first) if self.needs_input_grad[1] else None
        ws_bytes = _cached_ws_bytes(lib.fldr_corr81_bwd_workspace_bytes, B, C, H, W)
        ws = _workspace(ws_bytes, first.device)
        with _device_of(first):
            st = lib.fldr_corr81_bwd(_lib.ptr(first), _lib.strides(first), _lib.ptr(second), _lib.strides(second),
                                     _lib.ptr(gradOutput), _lib.strides(gradOutput), _lib.ptr(gradFirst),
                                     _lib.ptr(gradSecond), B, C, H, W, _lib.ptr(ws), ws_bytes, _stream_ptr(first.device))
        _lib.check(st)
        return gradFirst, gradSecond


def FunctionCorrelation(tensorFirst, tensorSecond):
    if not (torch.is_grad_enabled() and (tensorFirst.requires_grad or tensorSecond.requires_grad)):
        return _corr_forward(tensorFirst, tensorSecond)      # inference: nothing to record
    return _FunctionCorrelation.apply(tensorFirst, tensorSecond)


class ModuleCorrelation(torch.nn.Module):
    def __init__(self):
        super(ModuleCorrelation, self).__init__()

    def forward(self, tensorFirst, tensorSecond):
        return FunctionCorrelation(tensorFirst, tensorSecond)


def _corr_act_forward(tensorFirst, tensorSecond, negative_slope, out):
    B, C, H, W = tensorFirst.shape
    if out is None:
        out = torch.empty((B, 81, H, W), dtype=torch.float32, device=tensorFirst.device)
    lib = _lib.lib()
    with _device_of(tensorFirst):
        st = lib.fldr_corr81_fwd_act(_lib.ptr(tensorFirst), _lib.strides(tensorFirst), _lib.ptr(tensorSecond),
                                     _lib.strides(tensorSecond), _lib.ptr(out), out.stride(0), float(negative_slope),
                                     B, C, H, W, _stream_ptr(tensorFirst.device))
    _lib.check(st)
    return out


class _FunctionCorrelationLeakyReLU(torch.autograd.Function):
    """leaky_relu(correlation) with the activation fused in the forward's store epilogue; the backward scales the incoming
    gradient by the activation's slope (1 where the stored volume is positive, ``negative_slope`` elsewhere - the volume
    and its pre-activation have the same sign) and runs the correlation backward kernels (PWCNet.py:146-158 under autograd)."""

    @staticmethod
    def forward(ctx, first, second, negative_slope):
        out = _corr_act_forward(first, second, negative_slope, None)
        ctx.save_for_backward(first, second, out)
        ctx.slope = float(negative_slope)
        return out

    @staticmethod
    def backward(ctx, gradOutput):
        first, second, out = ctx.saved_tensors
        _check_cuda_f32("gradOutput", gradOutput)
        g = torch.where(out > 0, gradOutput, gradOutput * ctx.slope).contiguous()
        lib = _lib.lib()
        B, C, H, W = first.shape
        gradFirst = torch.empty_like(first) if ctx.needs_input_grad[0] else None
        gradSecond = torch.empty_like(first) if ctx.needs_input_grad[1] else None
        ws_bytes = _cached_ws_bytes(lib.fldr_corr81_bwd_workspace_bytes, B, C, H, W)
        ws = _workspace(ws_bytes, first.device)
        with _device_of(first):
            st = lib.fldr_corr81_bwd(_lib.ptr(first), _lib.strides(first), _lib.ptr(second), _lib.strides(second),
                                     _lib.ptr(g), _lib.strides(g), _lib.ptr(gradFirst), _lib.ptr(gradSecond), B, C, H, W,
                                     _lib.ptr(ws), ws_bytes, _stream_ptr(first.device))
        _lib.check(st)
        return gradFirst, gradSecond, None


def FunctionCorrelationLeakyReLU(tensorFirst, tensorSecond, negative_slope=0.1, out=None):
    """Next row: ``leaky_relu(FunctionCorrelation(first, second), negative_slope)`` as PWC-Net's
    decoder computes it (PWCNet.py:146-158), with the activation fused into the correlation's store epilogue.

    ``out``: optional preallocated contiguous buffer ``[B, Ctot >= 81, H, W]`` - the volume is written into
    ``out[:, :81]`` (the first block of ``torch.cat([tensorVolume, tensorFirst, tensorFlow, tensorFeat], 1)``,
    PWCNet.py:160) and that view is returned, so no separate concatenation copy of the volume is needed.
    Differentiable w.r.t. both feature maps (dense result only: ``out`` must be None when a gradient is needed)."""
    if not tensorFirst.is_cuda:
        raise NotImplementedError()
    _check_cuda_f32("first", tensorFirst)
    _check_cuda_f32("second", tensorSecond)
    assert (tensorFirst.is_contiguous() == True)
    assert (tensorSecond.is_contiguous() == True)
    assert tensorFirst.shape == tensorSecond.shape
    B, C, H, W = tensorFirst.shape
    if torch.is_grad_enabled() and (tensorFirst.requires_grad or tensorSecond.requires_grad):
        if out is not None:
            raise NotImplementedError("FunctionCorrelationLeakyReLU: writing into a caller's buffer is forward-only")
        return _FunctionCorrelationLeakyReLU.apply(tensorFirst, tensorSecond, float(negative_slope))
    if out is not None:
        _check_cuda_f32("out", out)
        assert out.is_contiguous() and out.shape[0] == B and out.shape[1] >= 81 and out.shape[2:] == (H, W)
    return _corr_act_forward(tensorFirst, tensorSecond, negative_slope, out)[:, :81]

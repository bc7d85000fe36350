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

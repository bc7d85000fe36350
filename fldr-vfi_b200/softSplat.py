"""Host mirror of the reference's ``softSplat.py`` over the sm_100a C-ABI library.

Same names, argument meaning and error behaviour as the reference (file:line = /root/reference/softSplat.py):

  Softsplat(strType='softmax').forward(img, flow, z=None)            355-361
  FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType)           320-352
  _FunctionSoftsplat.apply(input, flow)  (raw summation splat)       220-318

What differs (documented in DESIGN.md): one fused launch pair instead of ~12 torch kernels + a JIT string
kernel; no per-call regex templating (160-213); views are consumed through their strides instead of
``.contiguous()`` copies (231-232); non-fp32 tensors raise ``TypeError`` instead of being misread; a
non-finite flow skips the pixel instead of tripping a device assert (25-26).
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib


def _check_cuda_f32(name, t):
    if not t.is_cuda:
        raise NotImplementedError(f"{name}: CPU tensors are not supported (softSplat.py:251-252); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected torch.float32, got {t.dtype}")


def _stream_ptr(device):
    # raw handle of the current stream at CALL time (softSplat.py:246); the C accessor is ~20x cheaper than
    # torch.cuda.current_stream() and these ops are launch-latency bound on the small pyramid levels
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(device.index))


class _device_of:
    """Make ``t``'s device current for the raw launch (torch ops do this implicitly)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch._C._cuda_getDevice()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


_ws_cache = {}


def _workspace(nbytes, device):
    """Scratch for one call.  One grow-only buffer per (device, stream) is kept and reused: kernels of successive calls on a
    stream are ordered, so they can share it, and a 4K splat does not ask the allocator for 151 MB per call (calls issued on
    different streams get different buffers)."""
    nbytes = max(int(nbytes), 16)
    if torch.cuda.is_current_stream_capturing():          # graph capture: the graph's private pool owns its scratch
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, torch._C._cuda_getCurrentRawStream(device.index))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return buf


def release_workspaces():
    """Drop the cached scratch buffers (they are re-created on demand)."""
    _ws_cache.clear()


_ws_bytes_cache = {}


def _cached_ws_bytes(fn, *key):
    """*_workspace_bytes depends on the shape only; remember it per shape (one ctypes call less per op call)."""
    k = (fn.__name__,) + key
    v = _ws_bytes_cache.get(k)
    if v is None:
        v = _ws_bytes_cache[k] = int(fn(*key))
    return v


_CHECK_FLOW = os.environ.get("FLDR_B200_CHECK_FLOW", "0") not in ("", "0")
_flow_flags = {}


def _flow_flag(device):
    """Debug mode (FLDR_B200_CHECK_FLOW=1): the device word the splat kernels raise on a non-finite flow, one per device."""
    f = _flow_flags.get(device.index)
    if f is None:
        f = _flow_flags[device.index] = torch.zeros(1, dtype=torch.int32, device=device)
        with _device_of(f):
            _lib.check(_lib.lib().fldr_splat_set_nonfinite_flag(ctypes.c_void_p(f.data_ptr())))
    return f


def _check_flow_flag(dev):
    if _CHECK_FLOW:
        flag = _flow_flag(dev)
        if int(flag.item()):                               # synchronises: debug mode only
            flag.zero_()
            raise RuntimeError("non-finite flow: the reference asserts isfinite on the target coordinates (softSplat.py:25-26); "
                               "those pixels were skipped")


def _splat_forward(mode, tenInput, tenFlow, tenMetric, want_norm):
    N, C, H, W = tenInput.shape
    assert tenFlow.shape[1] == 2                      # softSplat.py:227
    assert tenFlow.shape[0] == N and tenFlow.shape[2] == H and tenFlow.shape[3] == W   # 228-229
    dev = tenInput.device
    if _CHECK_FLOW:
        _flow_flag(dev)
    ext = _lib.ext()
    if ext is not None:
        out, norm = ext.splat_fwd(mode, tenInput, tenFlow, tenMetric, bool(want_norm))
        _check_flow_flag(dev)
        return out, norm
    lib = _lib.lib()
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
    has_norm = mode in (1, 2, 3)
    norm = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev) if (want_norm and has_norm) else None
    metric = tenMetric
    if metric is not None:
        metric = metric.expand(N, 1, H, W)
    ws_bytes = _cached_ws_bytes(lib.fldr_splat_fwd_workspace_bytes, mode, N, C, H, W)
    ws = _workspace(ws_bytes, dev)
    with _device_of(tenInput):
        st = lib.fldr_splat_fwd(mode, _lib.ptr(tenInput), _lib.strides(tenInput), _lib.ptr(tenFlow),
                                _lib.strides(tenFlow), _lib.ptr(metric), None if metric is None else _lib.strides(metric),
                                _lib.ptr(out), _lib.ptr(norm), N, C, H, W, _lib.ptr(ws), ws_bytes, _stream_ptr(dev))
    _lib.check(st)
    _check_flow_flag(dev)
    return out, norm


def _splat_backward(mode, tenInput, tenFlow, tenMetric, out, norm, gradOutput, need):
    ext = _lib.ext()
    if ext is not None:
        return ext.splat_bwd(mode, tenInput, tenFlow, tenMetric, out, norm, gradOutput, bool(need[0]), bool(need[1]), bool(need[2]))
    lib = _lib.lib()
    N, C, H, W = tenInput.shape
    dev = tenInput.device
    gin = torch.empty((N, C, H, W), dtype=torch.float32, device=dev) if need[0] else None
    gfl = torch.empty((N, 2, H, W), dtype=torch.float32, device=dev) if need[1] else None
    gme = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev) if need[2] else None
    metric = tenMetric
    if metric is not None:
        metric = metric.expand(N, 1, H, W)
    ws_bytes = _cached_ws_bytes(lib.fldr_splat_bwd_workspace_bytes, mode, N, C, H, W)
    ws = _workspace(ws_bytes, dev)
    with _device_of(tenInput):
        st = lib.fldr_splat_bwd(mode, _lib.ptr(tenInput), _lib.strides(tenInput), _lib.ptr(tenFlow),
                                _lib.strides(tenFlow), _lib.ptr(metric), None if metric is None else _lib.strides(metric),
                                _lib.ptr(out), _lib.ptr(norm), _lib.ptr(gradOutput), _lib.strides(gradOutput),
                                _lib.ptr(gin), _lib.ptr(gfl), _lib.ptr(gme), N, C, H, W,
                                _lib.ptr(ws), ws_bytes, _stream_ptr(dev))
    _lib.check(st)
    return gin, gfl, gme


class _FunctionSoftsplat(torch.autograd.Function):
    """Raw summation splat (softSplat.py:220-318): ``apply(input, flow)`` -> ``S``; backward -> (gradInput, gradFlow)."""

    @staticmethod
    def forward(self, input, flow):
        _check_cuda_f32("input", input)
        _check_cuda_f32("flow", flow)
        out, _ = _splat_forward(_lib.SPLAT_MODES["raw"], input, flow, None, False)
        self.save_for_backward(input, flow)
        return out

    @staticmethod
    def backward(self, gradOutput):
        input, flow = self.saved_tensors
        _check_cuda_f32("gradOutput", gradOutput)
        need = (self.needs_input_grad[0], self.needs_input_grad[1], False)
        gin, gfl, _ = _splat_backward(_lib.SPLAT_MODES["raw"], input, flow, None, None, None, gradOutput, need)
        return gin, gfl


class _FusedSoftsplat(torch.autograd.Function):
    """``FunctionSoftsplat`` as one op: mode pre-processing, splat, normalise and post-scale fused."""

    @staticmethod
    def forward(self, tenInput, tenFlow, tenMetric, mode):
        need_bwd = any(self.needs_input_grad[:3])
        out, norm = _splat_forward(mode, tenInput, tenFlow, tenMetric, need_bwd)
        self.mode = mode
        if need_bwd:
            self.save_for_backward(tenInput, tenFlow, tenMetric, out, norm)
        return out

    @staticmethod
    def backward(self, gradOutput):
        tenInput, tenFlow, tenMetric, out, norm = self.saved_tensors
        _check_cuda_f32("gradOutput", gradOutput)
        need = (self.needs_input_grad[0], self.needs_input_grad[1], self.needs_input_grad[2] and tenMetric is not None)
        uses_metric = self.mode in (_lib.SPLAT_MODES["linear"], _lib.SPLAT_MODES["softmax"])
        need_k = (need[0], need[1], need[2] and uses_metric)
        gin, gfl, gme = _splat_backward(self.mode, tenInput, tenFlow, tenMetric if uses_metric else None, out, norm,
                                        gradOutput, need_k)
        if need[2] and not uses_metric:
            gme = torch.zeros_like(tenMetric)
        elif gme is not None and tenMetric is not None and gme.shape != tenMetric.shape:
            gme = gme.sum_to_size(tenMetric.shape)
        return gin, gfl, gme, None


def FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType):
    assert (tenMetric is None or tenMetric.shape[1] == 1)                       # softSplat.py:321
    assert (strType in ['summation', 'average', 'linear', 'softmax'])           # softSplat.py:322
    _check_cuda_f32("tenInput", tenInput)
    _check_cuda_f32("tenFlow", tenFlow)
    if tenMetric is not None:
        _check_cuda_f32("tenMetric", tenMetric)
    if strType == 'linear' and tenMetric is None:
        raise TypeError("strType 'linear' needs tenMetric (softSplat.py:328)")
    mode = _lib.SPLAT_MODES[strType]
    if not (torch.is_grad_enabled() and (tenInput.requires_grad or tenFlow.requires_grad
                                         or (tenMetric is not None and tenMetric.requires_grad))):
        # inference (main.py:test runs under no_grad): nothing to record, skip the autograd.Function machinery
        return _splat_forward(mode, tenInput, tenFlow, tenMetric, False)[0]
    return _FusedSoftsplat.apply(tenInput, tenFlow, tenMetric, mode)


class Softsplat(nn.Module):
    def __init__(self, strType='softmax'):
        super(Softsplat, self).__init__()
        self.strType = strType

    def forward(self, img, flow, z=None):
        return FunctionSoftsplat(img, flow, z, self.strType)

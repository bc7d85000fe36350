"""Torch-CPU restatement of the reference splat op (TEST INFRASTRUCTURE - see oracle/__init__.py).

Every function cites the lines of ``/root/reference/softSplat.py`` it restates.  Works in
fp32 (bit-faithful per-term arithmetic, different summation order) and fp64 (the arbiter
for the non-deterministic atomic order of the GPU kernels).
"""
import torch


def _corners(flow):
    """Target coordinates, integer corners and bilinear weights.

    Restates softSplat.py:23-38 (identical at 68-83 and 114-125): ``X = float(x) + u``,
    ``NW = floor``, the four weights built from the *integer corner minus X* products.
    Returns ``(X, Y, x0, y0)`` with x0/y0 as int64 tensors, all of shape [N, H, W].
    """
    N, _, H, W = flow.shape
    gx = torch.arange(W, dtype=flow.dtype).view(1, 1, W)
    gy = torch.arange(H, dtype=flow.dtype).view(1, H, 1)
    X = gx + flow[:, 0]
    Y = gy + flow[:, 1]
    assert bool(torch.isfinite(X).all()) and bool(torch.isfinite(Y).all())  # softSplat.py:25-26
    x0 = torch.floor(X)
    y0 = torch.floor(Y)
    return X, Y, x0, y0


def _weights(X, Y, x0, y0):
    """softSplat.py:35-38.  Order: NW, NE, SW, SE; each with its (dx, dy) corner offset."""
    x1 = x0 + 1
    y1 = y0 + 1
    wNW = (x1 - X) * (y1 - Y)
    wNE = (X - x0) * (y1 - Y)
    wSW = (x1 - X) * (Y - y0)
    wSE = (X - x0) * (Y - y0)
    return [(0, 0, wNW), (1, 0, wNE), (0, 1, wSW), (1, 1, wSE)]


def splat_raw(inp, flow):
    """Summation splat = ``kernel_Softsplat_updateOutput`` (softSplat.py:12-52).

    ``out[n,c,cy,cx] += in[n,c,y,x] * w`` for the 4 corners inside the frame (39-50);
    corners outside the frame are dropped.  Output zero-initialised (234).
    """
    N, C, H, W = inp.shape
    assert flow.shape == (N, 2, H, W)  # softSplat.py:227-229
    X, Y, x0, y0 = _corners(flow)
    out = torch.zeros(N, C, H * W, dtype=inp.dtype)
    src = inp.reshape(N, C, H * W)
    x0i = x0.long()
    y0i = y0.long()
    for dx, dy, w in _weights(X, Y, x0, y0):
        cx = x0i + dx
        cy = y0i + dy
        valid = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
        idx = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).reshape(N, 1, H * W).expand(N, C, H * W)
        contrib = src * w.reshape(N, 1, H * W)  # in * weight, softSplat.py:40
        contrib = torch.where(valid.reshape(N, 1, H * W), contrib, torch.zeros((), dtype=inp.dtype))
        out.scatter_add_(2, idx, contrib)
    return out.reshape(N, C, H, W)


def splat_raw_grad_input(flow, grad_out):
    """``kernel_Softsplat_updateGradInput`` (softSplat.py:54-98): bilinear gather of gradOutput."""
    N, C, H, W = grad_out.shape
    X, Y, x0, y0 = _corners(flow)
    g = grad_out.reshape(N, C, H * W)
    gin = torch.zeros(N, C, H * W, dtype=grad_out.dtype)
    x0i = x0.long()
    y0i = y0.long()
    for dx, dy, w in _weights(X, Y, x0, y0):
        cx = x0i + dx
        cy = y0i + dy
        valid = ((cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)).reshape(N, 1, H * W)
        idx = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).reshape(N, 1, H * W).expand(N, C, H * W)
        val = torch.gather(g, 2, idx) * w.reshape(N, 1, H * W)
        gin += torch.where(valid, val, torch.zeros((), dtype=g.dtype))
    return gin.reshape(N, C, H, W)


def splat_raw_grad_flow(inp, flow, grad_out):
    """``kernel_Softsplat_updateGradFlow`` (softSplat.py:100-158).

    Component 0 uses the d/dX weights (130-134), component 1 the d/dY weights (135-139);
    sum over channels of ``in * gradOut[corner] * dw`` (141-155).
    """
    N, C, H, W = inp.shape
    X, Y, x0, y0 = _corners(flow)
    x1 = x0 + 1
    y1 = y0 + 1
    one = torch.ones((), dtype=inp.dtype)
    dwx = [(0, 0, -one * (y1 - Y)), (1, 0, +one * (y1 - Y)), (0, 1, -one * (Y - y0)), (1, 1, +one * (Y - y0))]
    dwy = [(0, 0, (x1 - X) * -one), (1, 0, (X - x0) * -one), (0, 1, (x1 - X) * +one), (1, 1, (X - x0) * +one)]
    g = grad_out.reshape(N, C, H * W)
    src = inp.reshape(N, C, H * W)
    x0i = x0.long()
    y0i = y0.long()
    out = torch.zeros(N, 2, H * W, dtype=inp.dtype)
    for comp, dws in enumerate((dwx, dwy)):
        for dx, dy, dw in dws:
            cx = x0i + dx
            cy = y0i + dy
            valid = ((cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)).reshape(N, 1, H * W)
            idx = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).reshape(N, 1, H * W).expand(N, C, H * W)
            val = src * torch.gather(g, 2, idx) * dw.reshape(N, 1, H * W)
            val = torch.where(valid, val, torch.zeros((), dtype=g.dtype))
            out[:, comp] += val.sum(1)
    return out.reshape(N, 2, H, W)


class _SplatRaw(torch.autograd.Function):
    """Autograd pairing identical to ``_FunctionSoftsplat`` (softSplat.py:220-318)."""

    @staticmethod
    def forward(ctx, inp, flow):
        ctx.save_for_backward(inp, flow)
        return splat_raw(inp, flow)

    @staticmethod
    def backward(ctx, grad_out):
        inp, flow = ctx.saved_tensors
        gi = splat_raw_grad_input(flow, grad_out) if ctx.needs_input_grad[0] else None
        gf = splat_raw_grad_flow(inp, flow, grad_out) if ctx.needs_input_grad[1] else None
        return gi, gf


def function_softsplat(tenInput, tenFlow, tenMetric, strType, raw=None):
    """``FunctionSoftsplat`` (softSplat.py:320-352), this fork's semantics.

    * 'softmax' pre-scales ``(x+1)/2`` (334); ``tenMetric=None`` means weight 1 (335-336).
    * the normaliser's zeros are replaced by 1 in place (346) - holes come out as -1.
    * the final ``(y-0.5)*2`` is applied for EVERY mode (349).
    Differentiable (torch autograd through the restated ops + ``_SplatRaw``).  ``raw`` swaps the
    summation-splat autograd function (used by ``ref_host`` to plug the reference kernels in).
    """
    assert tenMetric is None or tenMetric.shape[1] == 1
    assert strType in ['summation', 'average', 'linear', 'softmax']
    if strType == 'average':
        tenInput = torch.cat([tenInput, tenInput.new_ones(tenInput.shape[0], 1, tenInput.shape[2], tenInput.shape[3])], 1)
    elif strType == 'linear':
        tenInput = torch.cat([tenInput * tenMetric, tenMetric], 1)
    elif strType == 'softmax':
        tenInput = (tenInput + 1) / 2
        if tenMetric is None:
            ones = tenInput.new_ones(tenInput.shape[0], 1, tenInput.shape[2], tenInput.shape[3])
            tenInput = torch.cat([tenInput * 1, ones], 1)
        else:
            tenInput = torch.cat([tenInput * tenMetric.exp(), tenMetric.exp()], 1)
    tenOutput = (raw or _SplatRaw.apply)(tenInput, tenFlow)
    if strType != 'summation':
        tenNormalize = tenOutput[:, -1:, :, :].clone()
        hole = tenNormalize == 0.0
        # in-place masked fill at softSplat.py:346: value 1, gradient blocked where masked
        tenNormalize = torch.where(hole, torch.ones((), dtype=tenOutput.dtype), tenNormalize)
        tenOutput = tenOutput[:, :-1, :, :] / tenNormalize
    tenOutput = (tenOutput - 0.5) * 2
    return tenOutput


def function_softsplat_grads(tenInput, tenFlow, tenMetric, strType, gradY):
    """Gradients of ``function_softsplat`` w.r.t. (input, flow, metric) via autograd of the
    restatement (metric grad has no reference kernel: it falls out of autograd through
    ``x*exp(z)`` / ``exp(z)``)."""
    x = tenInput.detach().clone().requires_grad_(True)
    f = tenFlow.detach().clone().requires_grad_(True)
    z = None if tenMetric is None else tenMetric.detach().clone().requires_grad_(True)
    y = function_softsplat(x, f, z, strType)
    wrt = [x, f] + ([z] if z is not None else [])
    grads = torch.autograd.grad(y, wrt, gradY, allow_unused=True)
    gz = grads[2] if z is not None else None
    return y.detach(), grads[0], grads[1], gz

"""Lift pieces of the staged, unmodified reference sources (``baseline/_ref/``, see fetch_ref.py) into callables, so the
"reference operator sequence" timed or compared by tools/ and tests/ is the reference's OWN text rather than a
restatement kept in this repository.  Works wherever baseline/_ref exists (the build container and the GPU box).

    bwarp = ref_src.bwarp(device)                 # DCTVFInet.bwarp, fLDRnet.py:546-581, as written
    blend = ref_src.blend()                       # the statements fLDRnet.py:510-524
    pwcb  = ref_src.pwc_backward()                # PWC-Net's Backward, OpticalFlow/PWCNet.py:116-143
"""
import ast
import os
import textwrap
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REFDIR, "fLDRnet.py")) and os.path.exists(os.path.join(REFDIR, "OpticalFlow", "PWCNet.py"))


def _function_source(path, name, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    scope = tree
    if cls is not None:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    fn = next(n for n in ast.walk(scope) if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))


class _CreateOnDevice:
    """``torch`` as seen by lifted code, with ``arange`` / ``ones`` allocating on ``device``: turns the reference's
    "build on the CPU, then .to(device)" into device-resident construction without touching its text."""

    def __init__(self, device):
        self._device = device

    def __getattr__(self, name):
        return getattr(torch, name)

    def arange(self, *a, **k):
        k.setdefault("device", self._device)
        return torch.arange(*a, **k)

    def ones(self, *a, **k):
        k.setdefault("device", self._device)
        return torch.ones(*a, **k)


def bwarp(device, create_on_device=False):
    """``f(x, flo, withmask=True)``: the reference's method bound to a stand-in ``self`` that only carries ``device``.
    ``create_on_device``: same text, but its mesh grid and ones tensor are allocated on the device (an idealised
    baseline: as written they are built on the CPU and copied over on every call, fLDRnet.py:555-569)."""
    ns = {"torch": _CreateOnDevice(device) if create_on_device else torch, "nn": torch.nn}
    exec(compile(_function_source(os.path.join(REFDIR, "fLDRnet.py"), "bwarp", "DCTVFInet"), "fLDRnet.py:bwarp", "exec"), ns)
    owner = types.SimpleNamespace(device=device)
    return lambda x, flo, withmask=True: ns["bwarp"](owner, x, flo, withmask=withmask)


def pwc_backward():
    """``f(tensorInput, tensorFlow)`` with fresh caches per call site (the reference keeps its grid / ones tensors in
    dicts keyed by shape)."""
    ns = {"torch": torch}
    exec(compile(_function_source(os.path.join(REFDIR, "OpticalFlow", "PWCNet.py"), "Backward"), "PWCNet.py:Backward", "exec"), ns)
    grid, ones = {}, {}
    return lambda tensorInput, tensorFlow: ns["Backward"](None, tensorInput, tensorFlow, grid, ones)


def to_pca_diff():
    """``f(im, params, args, mean, EV, mean_vec)``: the reference's own function text (pca_comp.py:473-528)."""
    import time
    ns = {"torch": torch, "nn": torch.nn, "time": time}
    exec(compile(_function_source(os.path.join(REFDIR, "pca_comp.py"), "to_pca_diff"), "pca_comp.py:to_pca_diff", "exec"), ns)
    return ns["to_pca_diff"]


def blend():
    """``f(refine_out, T_param, t_value, warped0, warped1, im0_tot, im1_tot, x_l) -> (out_l, occ_0_l)`` executing the
    reference's statements from ``num_softmax_combs = 6`` to ``out_l /=divisor``."""
    import torch.nn.functional as F
    lines = open(os.path.join(REFDIR, "fLDRnet.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if "num_softmax_combs = 6" in l)
    i1 = next(i for i, l in enumerate(lines) if i > i0 and "out_l /=divisor" in l)
    code = compile(textwrap.dedent("\n".join(l for l in lines[i0:i1 + 1] if l.strip())), "fLDRnet.py:510-524", "exec")

    def run(refine_out, T_param, t_value, warped0, warped1, im0_tot, im1_tot, x_l):
        ns = {"torch": torch, "F": F, "self": types.SimpleNamespace(T_param=T_param), "refine_out": refine_out, "t_value": t_value,
              "warped_img0_l": warped0, "warped_img1_l": warped1, "im0_tot": im0_tot, "im1_tot": im1_tot, "x_l": x_l}
        exec(code, ns)
        return ns["out_l"], ns["occ_0_l"]
    return run

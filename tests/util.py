"""Shared tolerances and fixture loading for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json north_star: softmax splat within 1e-4 abs / 1e-5 rel (atomic order is non-deterministic),
# correlation within 1e-5 relative in fp32.
SPLAT_ATOL = 1e-4
SPLAT_RTOL = 1e-5
CORR_RTOL = 1e-5


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def assert_splat_close(got, ref, what="", mag=None, cond=None):
    """|got-ref| <= atol*mag + rtol*|ref|.  ``mag`` (default max(1, max|ref|)) scales the absolute part for
    un-normalised quantities (raw sums, gradients) whose magnitude is not O(1); for the normalised splat
    output (|y| <= 1) it is 1 and the bound is exactly the north_star's 1e-4 abs / 1e-5 rel."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if mag is None:
        mag = max(1.0, float(ref.abs().max()))
    err = (got - ref).abs()
    bound = SPLAT_ATOL * mag + SPLAT_RTOL * ref.abs()
    if cond is not None:
        # ill-conditioned elements (gradients through a near-zero normaliser cancel terms ~1/norm): allow 10x the
        # oracle's own fp32-vs-fp64 discrepancy at that element on top of the flat bound
        bound = bound + 10.0 * cond.detach().double().cpu().abs()
    bad = err > bound
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements out of tolerance, max err {float(err.max()):.3e} (mag {mag:.3g})"


def assert_corr_close(got, ref, scale, what=""):
    """|got-ref| <= 1e-5 * scale + tiny, with scale = sum_c|f1*f2|/C per output element: '1e-5 relative' made
    well-posed for dot products that may cancel to ~0."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = scale.detach().double().cpu()
    err = (got - ref).abs()
    bound = CORR_RTOL * scale + 1e-7
    bad = err > bound
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements out of tolerance, max err {float(err.max()):.3e}, max ratio {float((err / (scale + 1e-30)).max()):.3e}"

"""End-to-end drop-in test: the untouched fLDRnet (reference files staged in baseline/_ref, shipped checkpoint) gives
the same interpolated frame with the sm_100a drop-ins as with its own CuPy kernels - PSNR within 0.01 dB (north_star c).
Reduced frame size to keep the GPU suite short; bench-scale 4K numbers live in profiles/ (baseline/e2e_fldrnet.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "fLDRnet.py")), reason="baseline/_ref not staged")]


def test_fldrnet_psnr_parity_with_dropins():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "e2e_fldrnet.py"), "--reps", "1",
                        "--height", "720", "--width", "1280"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["e2e_fldrnet"] == "ok"
    assert d["softSplat_module"]["ours"].endswith(os.path.join("dropin", "softSplat.py"))
    assert d["softSplat_module"]["reference"].endswith(os.path.join("_ref", "softSplat.py"))
    assert d["psnr_abs_diff_dB"] <= 0.01, d
    # with the bwarp method replaced by the fused gather kernel as well
    assert "fldr_vfi_b200.warp.bwarp" in d["softSplat_module"]["ours_warp"]
    assert d["with_bwarp_row"]["psnr_abs_diff_dB"] <= 0.01, d
    # ... and with torch.cuda.empty_cache() made a no-op on top (integrate.keep_allocator_cache): results unaffected
    assert d["with_bwarp_row_and_allocator_cache_kept"]["psnr_abs_diff_dB"] <= 0.01, d


def test_fldrnet_psnr_parity_4k():
    """north_star (c) at its literal size: 4096x2160 X-Test-shaped triplet, --papermodel --test5scales, PSNR of the
    interpolated frame within 0.01 dB of the run with the reference's own kernels - for the two-op drop-in and with the
    bwarp row replaced as well."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "e2e_fldrnet.py"), "--reps", "1",
                        "--variants", "reference,ours,ours_warp,ours_rows"], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert "unavailable" not in d, d
    assert d["frame"] == "4096x2160"
    for tag in ("ours", "ours_warp", "ours_rows"):
        assert d["variants"][tag]["psnr_abs_diff_dB"] <= 0.01, d
    assert "pca.to_pca_diff" in d["variants"]["ours_rows"]["ops"]

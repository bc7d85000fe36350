"""Stage the reference files the GPU baseline needs into ``baseline/_ref/`` (git-ignored; it travels to the GPU box
with the snapshot, where /root/reference does not exist).  Run in the build container:  python baseline/fetch_ref.py

Nothing is modified and nothing lands in git history; the files are used to run the reference's own CuPy kernels on
the same B200 (through ``baseline/cupy_shim``) as the reported baseline and as a second GPU oracle.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FLDR_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["softSplat.py", os.path.join("OpticalFlow", "correlation.py"),
         # end-to-end harness (baseline/e2e_fldrnet.py): the untouched model + its runner + the shipped checkpoint
         "fLDRnet.py", "useful.py", "pca_comp.py", "utils.py", "inter4kreader.py", "run_on_your_images.py",
         os.path.join("OpticalFlow", "PWCNet.py"),
         os.path.join("checkpoint_dir", "fLDRnet_X4K1000FPS_exp1", "fLDRnet_X4K1000FPS_exp1_best_PSNR.pt")]


def fetch(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"[baseline/fetch_ref] {REF} not present; keeping {DST} as is")
        return os.path.isdir(DST)
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    if verbose:
        print(f"[baseline/fetch_ref] staged {len(FILES)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)

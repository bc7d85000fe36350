"""Torch-CPU restatement of the block-PCA feature extraction (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates ``/root/reference/pca_comp.py:473-528`` (``to_pca_diff``), the first device step of every fLDRnet forward
(``fLDRnet.py:146``): 8x8 unfold of the two stacked frames, subtraction of the block mean vector, projection on the first
16 eigenvectors in float64, optional division by ``mean_vec``, reorder to ``[chan * 16, H/8, W/8]``, global min/max
normalisation to [-1, 1].  Pinned against ``tests/golden/pca_*.npz``, which ``tests/golden/make_golden.py --pca-only``
produced by executing the reference's own function text.
"""
import torch


def to_pca_diff(im, mean, EV, mean_vec=None, wiS=8):
    """im [chan, H, W] float32; mean [wiS*wiS] float64; EV [ncomp, wiS*wiS] float64; mean_vec [ncomp] float64 or None
    (``args.mean_vector_norm`` off).  Returns float64 ``[chan * ncomp, H/wiS, W/wiS]``."""
    chan, H, W = im.shape
    if H % wiS or W % wiS:
        raise Exception("in to_pca_diff the image is not padded right." + str(H) + " " + str(W))       # pca_comp.py:486-487
    by, bx = H // wiS, W // wiS
    ncomp = EV.shape[0]
    # pca_comp.py:489-499: Unfold + the reshape / permute chain = rows ordered (chan, block_x, block_y), 64-vector = (ky, kx)
    blocks = im.reshape(chan, by, wiS, bx, wiS).permute(0, 3, 1, 2, 4).reshape(chan * bx * by, wiS * wiS)
    loc = blocks.to(torch.float64) - mean                                                              # 502
    t = torch.matmul(loc, EV.permute(1, 0))                                                            # 507
    if mean_vec is not None:
        t = t / mean_vec                                                                               # 510-511
    t = t.reshape(chan, bx, by, ncomp).permute(0, 3, 2, 1).reshape(-1, by, bx)                          # 516-518
    mi, ma = torch.min(t), torch.max(t)                                                                # 521-522
    return ((t - mi) / (ma - mi)) * 2 - 1                                                              # 523-526

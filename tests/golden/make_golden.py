"""Generate the golden fixtures in this directory from the REFERENCE'S OWN KERNEL TEXT.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Each ``*.npz`` holds seeded inputs and the outputs of the reference kernels executed on the
host through ``oracle/_ref/libref_host.so`` (see oracle/build_ref.py).  Splat outputs depend
on the atomic summation order; the host run is sequential-per-thread, so values are one
valid order - consumers compare with the tolerances of BASELINE.json (1e-4 abs / 1e-5 rel),
the correlation with 1e-5 relative to sum|f1*f2|/C.
``wrapper_*`` entries went through ``FunctionSoftsplat``'s torch glue as RESTATED in
oracle/splat_oracle.py (softSplat.py:320-352) around the reference kernels.
"""
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_host, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def splat_case(name, N, C, H, W, regime, scale, seed):
    x = synth.features(N, C, H, W, seed=seed)
    fl = synth.flow(N, H, W, regime, seed=seed + 1) * scale
    z = synth.metric(N, H, W, seed=seed + 2)
    g = synth.grad((N, C, H, W), seed=seed + 3)
    out = {"input": x, "flow": fl, "metric": z, "grad_out": g}
    out["raw_out"] = ref_host.splat_update_output(x, fl)
    out["raw_grad_input"] = ref_host.splat_update_grad_input(x, fl, g)
    out["raw_grad_flow"] = ref_host.splat_update_grad_flow(x, fl, g)
    for mode in ["summation", "average", "linear", "softmax"]:
        xi = x.clone().requires_grad_(True)
        fi = fl.clone().requires_grad_(True)
        zi = z.clone().requires_grad_(True)
        m = None if mode in ("summation", "average") else zi
        y = ref_host.function_softsplat(xi, fi, m, mode)
        wrt = [xi, fi] + ([zi] if m is not None else [])
        grads = torch.autograd.grad(y, wrt, g)
        out[f"wrapper_{mode}_out"] = y.detach()
        out[f"wrapper_{mode}_grad_input"] = grads[0]
        out[f"wrapper_{mode}_grad_flow"] = grads[1]
        if m is not None:
            out[f"wrapper_{mode}_grad_metric"] = grads[2]
    # softmax with metric=None (feature-splat path, fLDRnet.py:386-387)
    xi = x.clone().requires_grad_(True)
    fi = fl.clone().requires_grad_(True)
    y = ref_host.function_softsplat(xi, fi, None, "softmax")
    gi, gf = torch.autograd.grad(y, [xi, fi], g)
    out["wrapper_softmax_nometric_out"] = y.detach()
    out["wrapper_softmax_nometric_grad_input"] = gi
    out["wrapper_softmax_nometric_grad_flow"] = gf
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.numpy() for k, v in out.items()})
    print(name, {k: tuple(v.shape) for k, v in out.items() if k.startswith("raw")})


def corr_case(name, B, C, H, W, seed):
    f1 = synth.features(B, C, H, W, seed=seed)
    f2 = synth.features(B, C, H, W, seed=seed + 1)
    out = ref_host.corr_update_output(f1, f2)
    g = synth.grad(tuple(out.shape), seed=seed + 2)
    g1, g2 = ref_host.corr_update_grads(f1, f2, g)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), first=f1.numpy(), second=f2.numpy(), out=out.numpy(),
                        grad_out=g.numpy(), grad_first=g1.numpy(), grad_second=g2.numpy(),
                        rbot0=ref_host.corr_rearrange(f1).numpy())
    print(name, tuple(out.shape))


def _reference_bwarp():
    """The reference's own ``bwarp`` method (fLDRnet.py:546-581), lifted out of its class by ast and run on the CPU.
    (Importing fLDRnet.py whole needs cupy and a GPU; the method itself is plain torch.)"""
    import ast
    import textwrap
    import types
    src = open("/root/reference/fLDRnet.py").read()
    tree = ast.parse(src)
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "DCTVFInet"
              for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "bwarp")
    code = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(code, "fLDRnet.py:bwarp", "exec"), ns)
    owner = types.SimpleNamespace(device=torch.device("cpu"))
    return lambda x, flo, withmask=True: ns["bwarp"](owner, x, flo, withmask=withmask)


def warp_case(name, N, C, H, W, regime, scale, seed, alpha=-1.894):
    bwarp = _reference_bwarp()
    x0 = synth.image(N, C, H, W, seed=seed)
    x1 = synth.image(N, C, H, W, seed=seed + 1)
    fl = synth.flow(N, H, W, regime, seed=seed + 2) * scale
    out = {"ref": x0, "src": x1, "flow": fl, "alpha": torch.tensor(alpha)}
    with torch.no_grad():
        out["bwarp_mask"] = bwarp(x1, fl.clone(), True)
        out["bwarp_nomask"] = bwarp(x1, fl.clone(), False)
        z_alpha = torch.tensor([alpha, alpha], dtype=torch.float64)          # fLDRnet.py:360: a double Parameter
        out["metric"] = torch.mean(z_alpha[0] * torch.abs(x0 - out["bwarp_mask"]), dim=1, keepdim=True)   # fLDRnet.py:443
    assert out["metric"].dtype == torch.float32
    # gradients of the reference's own bwarp by autograd (grid_sample backward on the CPU)
    g = synth.grad((N, C, H, W), seed=seed + 3)
    xi, fi = x1.clone().requires_grad_(True), fl.clone().requires_grad_(True)
    gx, gf = torch.autograd.grad(bwarp(xi, fi, True), [xi, fi], g)
    out["grad_out"], out["grad_src"], out["grad_flow"] = g, gx, gf
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.numpy() for k, v in out.items()})
    print(name, tuple(out["bwarp_mask"].shape), "masked px", int((out["bwarp_mask"].abs().sum(1) == 0).sum()))


def pwcwarp_case(name, N, C, H, W, regime, scale, seed):
    """PWC-Net's own ``Backward`` source (OpticalFlow/PWCNet.py:116-143), lifted by ast.  The only edit is dropping the
    ``.cuda()`` of line 130: the reference builds its linspace grid on the CPU and then moves it, so the CPU values are
    the ones that count."""
    import ast
    import textwrap
    src = open("/root/reference/OpticalFlow/PWCNet.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "Backward")
    code = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])).replace(".cuda()", "")
    ns = {"torch": torch}
    exec(compile(code, "PWCNet.py:Backward", "exec"), ns)
    x = synth.features(N, C, H, W, seed=seed)
    fl = synth.flow(N, H, W, regime, seed=seed + 1) * scale
    with torch.no_grad():
        out = ns["Backward"](None, x, fl.clone(), {}, {})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), input=x.numpy(), flow=fl.numpy(), out=out.numpy())
    print(name, tuple(out.shape), "masked px", int((out.abs().sum(1) == 0).sum()))


def _reference_blend_lines():
    """fLDRnet.py's own source lines from ``num_softmax_combs = 6`` to ``out_l /=divisor`` (510-524), dedented."""
    import textwrap
    lines = open("/root/reference/fLDRnet.py").read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if "num_softmax_combs = 6" in l)
    i1 = next(i for i, l in enumerate(lines) if i > i0 and "out_l /=divisor" in l)
    return textwrap.dedent("\n".join(l for l in lines[i0:i1 + 1] if l.strip()))


def blend_case(name, N, C, H, W, seed, temperature=1.0, t=0.5):
    import types
    import torch.nn.functional as F
    code = _reference_blend_lines()
    imgs = [synth.image(N, C, H, W, seed=seed + k) for k in range(6)]
    refine_out = synth.grad((N, 8, H, W), seed=seed + 10) * 3.0          # 6 logits + 2 spare channels, like the U-Net output
    t_value = torch.full((N, 1), t) + 0.1 * torch.arange(N, dtype=torch.float32).view(N, 1) / max(N, 1)
    x_l = torch.stack([imgs[4], imgs[5]], 2)                             # [N,C,2,H,W] as in the model
    owner = types.SimpleNamespace(T_param=torch.full((1,), temperature, dtype=torch.float64))     # fLDRnet.py:357
    ns = {"torch": torch, "F": F, "self": owner, "refine_out": refine_out, "t_value": t_value.view(N, 1, 1, 1),
          "warped_img0_l": imgs[0], "warped_img1_l": imgs[1], "im0_tot": imgs[2], "im1_tot": imgs[3], "x_l": x_l}
    with torch.no_grad():
        exec(compile(code, "fLDRnet.py:510-524", "exec"), ns)
    assert ns["out_l"].dtype == torch.float64
    np.savez_compressed(os.path.join(HERE, name + ".npz"), refine_out=refine_out.numpy(), t_value=t_value.numpy(),
                        temperature=np.float64(temperature), **{f"img{k}": imgs[k].numpy() for k in range(6)},
                        out=ns["out_l"].numpy(), occ0=ns["occ_0_l"].numpy())
    print(name, tuple(ns["out_l"].shape), ns["out_l"].dtype)


def pca_case(name, chan, H, W, seed, mean_vector_norm):
    """pca_comp.py's own ``to_pca_diff`` (473-528), lifted by ast and executed on the CPU in float64 as the model does
    (fLDRnet.py:146: float32 frames, float64 mean / eigenvectors / mean_vec parameters)."""
    import ast
    import types
    import torch.nn as nn
    src = open("/root/reference/pca_comp.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "to_pca_diff")
    code = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
    import time
    ns = {"torch": torch, "nn": nn, "time": time}
    exec(compile(code, "pca_comp.py:to_pca_diff", "exec"), ns)
    g = torch.Generator().manual_seed(seed)
    im = synth.image(1, chan, H, W, seed=seed)[0]                                 # [chan, H, W] float32 in [-1, 1]
    mean = torch.randn(64, generator=g, dtype=torch.float64) * 0.1
    EV = torch.linalg.qr(torch.randn(64, 64, generator=g, dtype=torch.float64))[0][:16].contiguous()   # orthonormal rows
    mean_vec = torch.rand(16, generator=g, dtype=torch.float64) + 0.5
    params = types.SimpleNamespace(wiS=8, weightMat=None, components_fraction=0.25)
    args = types.SimpleNamespace(gpu="cpu", mean_vector_norm=mean_vector_norm)
    with torch.no_grad():
        out = ns["to_pca_diff"](im, params, args, mean, EV, mean_vec)
    assert out.dtype == torch.float64 and tuple(out.shape) == (chan * 16, H // 8, W // 8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), im=im.numpy(), mean=mean.numpy(), EV=EV.numpy(), mean_vec=mean_vec.numpy(),
                        mean_vector_norm=np.int32(mean_vector_norm), out=out.numpy())
    print(name, tuple(out.shape), out.dtype, float(out.min()), float(out.max()))


def pyramid_case(name, B, T, H, W, scales, n_levels, seed, align_corners=False, noise=False):
    """main.py's own list comprehension building ``input_gpu`` in ``test()`` (855-856), lifted by ast and evaluated on the
    CPU (``device = 'cpu'``) - ``F.interpolate(..., mode='bicubic')`` of the full-resolution frames per level."""
    import ast
    import types
    import torch.nn.functional as F
    src = open("/root/reference/main.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "test")
    stmt = next(n for n in ast.walk(fn) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name)
                and n.targets[0].id == "input_gpu" and isinstance(n.value, ast.ListComp))
    code = textwrap.dedent(ast.get_source_segment(src, stmt))
    C = 3
    if noise:
        frames = torch.randn(B, C, T, H, W, generator=torch.Generator().manual_seed(seed))     # white noise: every tap matters
    else:
        frames = synth.image(B, C * T, H, W, seed=seed).reshape(B, C, T, H, W).contiguous()
    args = types.SimpleNamespace(scales=list(scales), S_tst=n_levels, align_cornerse=align_corners)
    ns = {"torch": torch, "F": F, "args": args, "device": "cpu", "input_frames": frames, "B": B, "C": C, "T": T, "H": H, "W": W}
    with torch.no_grad():
        exec(compile(code, "main.py:855-856", "exec"), ns)
    levels = ns["input_gpu"]
    assert len(levels) == n_levels + 1 and levels[0].shape == frames.shape
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=frames.numpy(), scales=np.array(scales, dtype=np.int64),
                        n_levels=np.int32(n_levels), align_corners=np.int32(align_corners),
                        **{f"level{i}": levels[i].contiguous().numpy() for i in range(1, n_levels + 1)})
    print(name, [tuple(l.shape) for l in levels])


if __name__ == "__main__":
    if "--pyramid-only" in sys.argv:
        pyramid_case("pyramid_5lv", 1, 2, 64, 256, [8, 16, 32, 64, 128, 256], 5, 510)            # --test5scales: factors 1/2 .. 1/32
        pyramid_case("pyramid_noise_b2", 2, 2, 32, 160, [8, 16, 32, 64], 3, 520, noise=True)     # B = 2, W not a multiple of 128
        pyramid_case("pyramid_ac", 1, 2, 24, 40, [8, 16, 32], 2, 530, align_corners=True)        # --align_cornerse
        pyramid_case("pyramid_odd", 1, 2, 20, 36, [8, 24, 40], 2, 540, noise=True)               # factors 1/3, 1/5: generic taps
        sys.exit(0)
    if "--pca-only" in sys.argv:
        pca_case("pca_6x32x48", 6, 32, 48, 410, True)          # one sample (two RGB frames), mean-vector normalisation on
        pca_case("pca_12x16x24_nomv", 12, 16, 24, 420, False)  # B = 2, no mean-vector normalisation
        pca_case("pca_6x8x8", 6, 8, 8, 430, True)              # a single block per channel
        sys.exit(0)
    if "--blend-only" in sys.argv:
        blend_case("blend_t05", 2, 3, 16, 24, 210)                       # t = 0.5 / 0.55, T = 1 (the shipped checkpoint)
        blend_case("blend_temp", 1, 3, 9, 13, 220, temperature=0.37, t=0.25)   # odd sizes, learned temperature, t != 0.5
        blend_case("blend_c1", 3, 1, 8, 8, 230, temperature=2.0, t=0.8)
        sys.exit(0)
    if "--pwcwarp-only" in sys.argv:
        pwcwarp_case("pwcwarp_smooth", 2, 5, 24, 40, "F1", 10.0, 310)
        pwcwarp_case("pwcwarp_scatter", 1, 3, 17, 23, "F2", 1.0, 320)
        pwcwarp_case("pwcwarp_border", 2, 4, 20, 64, "FB", 1.0, 330)
        sys.exit(0)
    if "--warp-only" in sys.argv:
        warp_case("warp_smooth", 2, 3, 24, 40, "F1", 40.0, 110)     # image-like, smooth large flow
        warp_case("warp_scatter", 1, 2, 17, 23, "F2", 1.0, 120)     # C=2 (a flow warped by a flow, fLDRnet.py:474), odd sizes
        warp_case("warp_border", 2, 3, 20, 32, "FB", 1.0, 130)      # a quarter of the samples leave the frame
        warp_case("warp_identity", 1, 3, 8, 12, "F0", 1.0, 140)     # zero flow is NOT the identity (W/(W-1) quirk)
        sys.exit(0)
    splat_case("splat_smooth", 2, 3, 24, 40, "F1", 40.0, 10)    # image-like C=3, smooth large flow
    splat_case("splat_scatter", 1, 5, 17, 23, "F2", 1.0, 20)    # odd sizes, iid flow, C not multiple of 4
    splat_case("splat_converge", 1, 3, 16, 24, "F3", 1.0, 30)   # max contention
    splat_case("splat_border", 2, 4, 20, 32, "FB", 1.0, 40)     # ~25 % of targets out of frame
    splat_case("splat_identity", 1, 3, 8, 12, "F0", 1.0, 50)    # zero flow
    corr_case("corr_c40", 2, 40, 12, 20, 60)                    # C not multiple of 32
    corr_case("corr_c32_odd", 1, 32, 9, 11, 70)                 # odd H, W
    corr_case("corr_c196_tiny", 2, 196, 5, 8, 80)               # cfg2-literal level 6 shape

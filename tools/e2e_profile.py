"""Where does the untouched fLDRnet's 4K forward spend its time once splat, correlation and bwarp are replaced?
torch.profiler over one forward of baseline/e2e_fldrnet.py's `ours_warp` configuration.  python tools/e2e_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import e2e_fldrnet as E   # noqa: E402


def main():
    import torch
    sys.path.insert(0, os.path.join(ROOT, "fldr-vfi_b200", "dropin"))
    sys.path.insert(1, E.REFDIR)
    sys.path.append(os.path.join(E.HERE, "cupy_shim"))
    sys.path.append(os.path.join(E.HERE, "stubs"))
    os.chdir(E.REFDIR)
    _load = torch.load
    torch.load = lambda *a, **k: _load(*a, **{**k, "weights_only": k.get("weights_only", False)})
    import warnings
    warnings.simplefilter("ignore")
    sys.argv = ["run_on_your_images.py"]
    import run_on_your_images as R
    import fLDRnet
    import torch.nn.functional as F
    sys.path.insert(0, ROOT)
    from fldr_vfi_b200.integrate import patch_bwarp
    patch_bwarp(fLDRnet)
    model_net, device, args = R.prepare_model()
    model_net.eval()
    frames = E.synthetic_triplet(2160, 4096)
    t_value = torch.tensor([[0.5]])
    with torch.no_grad():
        input_frames = frames[:, :, :-1]
        B, C, T, H, W = input_frames.size()
        input_frames = input_frames.reshape(B, -1, H, W)
        div_pad = (2 ** args.S_tst) * 8
        Hp, Wp = (div_pad - H % div_pad) % div_pad, (div_pad - W % div_pad) % div_pad
        input_frames = F.pad(input_frames, (0, Wp, 0, Hp), args.padding).reshape(B, C, T, H + Hp, W + Wp)
        B, C, T, H, W = input_frames.shape
        input_gpu = [F.interpolate(input_frames.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W), scale_factor=args.scales[0] / args.scales[i],
                                   mode="bicubic", align_corners=args.align_cornerse).to(device)
                     .reshape(B, T, C, int(H * (args.scales[0] / args.scales[i])), int(W * (args.scales[0] / args.scales[i]))).permute(0, 2, 1, 3, 4)
                     if i != 0 else input_frames.to(device) for i in range(args.S_tst + 1)]
        t_dev = t_value.to(device)

        def fwd():
            lst = [torch.zeros((B, int(args.img_ch * 2 * (8 ** 2) * 0.25), H // 8, W // 8), device=device) for _ in range(6)]
            out = model_net(lst, t_dev, normInput=[im.clone() for im in input_gpu], is_training=False, validation=False)
            torch.cuda.synchronize()
            return out
        fwd(); fwd()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            fwd()
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=48))
        print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=18, max_name_column_width=48))


if __name__ == "__main__":
    main()

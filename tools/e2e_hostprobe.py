"""Where does the host time of the untouched fLDRnet 4K forward go?  Wall time (no added synchronisation) accumulated inside
bwarp / softsplat / correlation / torch.cuda.empty_cache for the reference ops and for the drop-ins.
    python tools/e2e_hostprobe.py reference|ours"""
import os, sys, time, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import e2e_fldrnet as E
impl = sys.argv[1]
acc = collections.defaultdict(float); cnt = collections.Counter()


def wrap(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] += time.perf_counter() - t0; cnt[name] += 1
    return w


import torch
if impl.startswith("ours"):
    sys.path.insert(0, os.path.join(ROOT, "fldr-vfi_b200", "dropin"))
sys.path.insert(1, E.REFDIR)
sys.path.append(os.path.join(E.HERE, "cupy_shim")); sys.path.append(os.path.join(E.HERE, "stubs"))
os.chdir(E.REFDIR)
_load = torch.load
torch.load = lambda *a, **k: _load(*a, **{**k, "weights_only": k.get("weights_only", False)})
import warnings; warnings.simplefilter("ignore")
sys.argv = ["run_on_your_images.py"]
import run_on_your_images as R
import softSplat, fLDRnet
import torch.nn.functional as F
import OpticalFlow.correlation as corr
fLDRnet.DCTVFInet.bwarp = wrap("bwarp", fLDRnet.DCTVFInet.bwarp)
softSplat.FunctionSoftsplat = wrap("softsplat", softSplat.FunctionSoftsplat)
softSplat.Softsplat.forward = wrap("Softsplat.forward", softSplat.Softsplat.forward)
corr.FunctionCorrelation = wrap("correlation", corr.FunctionCorrelation)
torch.cuda.empty_cache = wrap("empty_cache", torch.cuda.empty_cache)
fLDRnet.to_pca_diff = wrap("to_pca_diff", fLDRnet.to_pca_diff) if hasattr(fLDRnet, "to_pca_diff") else None
model_net, device, args = R.prepare_model(); model_net.eval()
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
    input_gpu = [F.interpolate(input_frames.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W), scale_factor=args.scales[0] / args.scales[i], mode="bicubic",
                               align_corners=args.align_cornerse).to(device).reshape(B, T, C, int(H * (args.scales[0] / args.scales[i])), int(W * (args.scales[0] / args.scales[i]))).permute(0, 2, 1, 3, 4)
                 if i != 0 else input_frames.to(device) for i in range(args.S_tst + 1)]
    t_dev = t_value.to(device)
    for rep in range(4):
        lst = [torch.zeros((B, int(args.img_ch * 2 * 64 * 0.25), H // 8, W // 8), device=device) for _ in range(6)]
        torch.cuda.synchronize(); acc.clear(); cnt.clear()
        t0 = time.perf_counter()
        model_net(lst, t_dev, normInput=[im.clone() for im in input_gpu], is_training=False, validation=False)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(json.dumps({"impl": impl, "rep": rep, "forward_host_s": round(t1 - t0, 4), "drain_s": round(t2 - t1, 4),
                          "inside": {k: [round(v, 4), cnt[k]] for k, v in acc.items()}}), flush=True)

#!/usr/bin/env python
"""bench.py - fLDR-VFI custom-kernel hot path on B200: 4K frame-pairs/sec (splat + correlation).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = the hot path of ONE synthetic 4K frame pair (BASELINE.json configs[2]'s splat set + configs[1]'s
correlation pyramid at native 4K):
  * 12 softmax splats of a `--papermodel --test5scales` interpolation (fLDRnet.py:386-387,449-450):
    2 x image splat C=3 + metric at 2304x4096, 2 x feature splat C=48 (metric=None) at each of 288x512 ... 18x32
  * 5 correlation cost volumes of a PWC-Net pass on the pair (bidirectional, B=2, useful.py:114):
    (C,H,W) = (196,34,64) (128,68,128) (96,136,256) (64,272,512) (32,544,1024)
Frame pairs are independent: rank r processes its own pairs, no collective on the data path (weak scaling).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same step through the public
drop-in API from pinned HOST buffers (H2D of every input, D2H of every result inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H4K, W4K = 2304, 4096
FEATURE_LEVELS = [(288, 512), (144, 256), (72, 128), (36, 64), (18, 32)]
CORR_LEVELS = [(196, 34, 64), (128, 68, 128), (96, 136, 256), (64, 272, 512), (32, 544, 1024)]
METRIC = "4K frame-pairs/sec (splat+corr)"


def splat_bytes(N, C, H, W, metric):
    return 4 * N * H * W * (2 * C + 2 + (1 if metric else 0))


def corr_bytes(B, C, H, W):
    return 4 * B * H * W * (2 * C + 81)


def corr_flops(B, C, H, W):
    return 162 * C * B * H * W


def shared_config():
    """`config` of the JSON line - the SAME dict in the `ours` and `reference` arms (arm-specific notes go in `notes`)."""
    return {"workload": "one 4K frame pair per step per GPU: 12 softmax splats of fLDRnet --papermodel --test5scales "
                        "(2x image C=3+metric 2304x4096, 2x feature C=48 at 5 levels) + 5-level PWC correlation "
                        "pyramid at native 4K (B=2, C=196..32)",
            "flow_regime": "F1 smooth",
            "algorithmic_MB_per_step": 1820.2,
            "sharding": "frame pairs across ranks, no collective"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(call):
    """DRAM bytes per call of `call` from the committed ncu --set full capture (profiles/r2_traffic.json, else r1), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return int(json.load(open(os.path.join(ROOT, "profiles", name)))[call]["bytes"])
        except Exception:
            continue
    return None


# ----------------------------------------------------------------------------------------------- inputs
def make_pair_inputs(seed, frac=1.0):
    """Synthetic inputs of one 4K frame pair on the CPU: images seed+0, flow F1 seed+1,
    metric seed+2, features seed+3.  ``frac`` < 1 keeps the top ``frac`` of the rows of every tensor (bounded
    CPU sample for the reference arm)."""
    from oracle import synth   # input generator only - no reference arithmetic
    def rows(h):
        return max(1, int(round(h * frac)))
    calls = []
    H = rows(H4K)
    for d in range(2):
        calls.append(("splat_image", dict(
            x=synth.image(1, 3, H, W4K, seed=seed + d), flow=synth.flow(1, H, W4K, "F1", seed=seed + 10 + d),
            z=synth.metric(1, H, W4K, seed=seed + 20 + d))))
    for li, (h, w) in enumerate(FEATURE_LEVELS):
        h = rows(h)
        for d in range(2):
            calls.append((f"splat_feat_L{li}", dict(
                x=synth.features(1, 48, h, w, seed=seed + 30 + 2 * li + d),
                flow=synth.flow(1, h, w, "F1", seed=seed + 50 + 2 * li + d) * 8.0, z=None)))
    for (c, h, w) in CORR_LEVELS:
        h = rows(h)
        calls.append((f"corr_C{c}", dict(f1=synth.features(2, c, h, w, seed=seed + 70 + c),
                                         f2=synth.features(2, c, h, w, seed=seed + 71 + c))))
    return calls


def call_bytes(name, t):
    if name.startswith("splat"):
        N, C, H, W = t["x"].shape
        return splat_bytes(N, C, H, W, t["z"] is not None)
    B, C, H, W = t["f1"].shape
    return corr_bytes(B, C, H, W)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of this rank's GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def pin_rank_to_cores(local_rank, world):
    """Give every rank its own slice of the host cores nearest its GPU (NVML affinity mask), so the pinned staging buffers
    it allocates afterwards are first-touched - and the copy engines are fed - from that NUMA node.  Returns the cores."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = [i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1]
    except Exception:
        near = []
    allowed = sorted(os.sched_getaffinity(0))
    near = [c for c in near if c in allowed] or allowed
    per = max(1, len(near) // max(world, 1))
    mine = near[(local_rank * per) % len(near):][:per] or near
    try:
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), 8)))
    except Exception:
        pass
    return mine


def median_event_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    cores = pin_rank_to_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import fldr_vfi_b200._lib as L
    import fldr_vfi_b200.correlation as C
    import fldr_vfi_b200.softSplat as S
    L.lib()
    splat = S.Softsplat()

    host = make_pair_inputs(seed=1000 * rank)
    pinned = [(n, {k: (None if v is None else v.pin_memory()) for k, v in t.items()}) for n, t in host]
    devin = [(n, {k: (None if v is None else v.to(dev)) for k, v in t.items()}) for n, t in pinned]
    names = [n for n, _ in devin]
    step_bytes = sum(call_bytes(n, t) for n, t in host)
    h2d_bytes = sum(v.numel() * 4 for _, t in host for v in t.values() if v is not None)

    def run_call(name, t):
        if name.startswith("splat"):
            return splat(t["x"], t["flow"], t["z"])
        return C.FunctionCorrelation(tensorFirst=t["f1"], tensorSecond=t["f2"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # our kernels launched per step (driver memsets not counted): splat = zero fill + scatter + normalise, or ONE cooperative
    # kernel for frames with <= 40000 accumulator float4s (fldr_set_option "splat_fused_max"); correlation = 1
    def kernels_of(n, t):
        if not n.startswith("splat"):
            return 1
        N, C, H, W = t["x"].shape
        return 1 if N * ((C + 1 + 3) // 4) * H * W <= 40000 else 3
    launches_per_step = sum(kernels_of(n, t) for n, t in host)

    with torch.no_grad():
        for _ in range(args.warmup):
            for n, t in devin:
                run_call(n, t)

        # clocks / throttle reasons are sampled from here to the end of the e2e pass (every phase in between is under
        # load; the timed value region alone lasts only tens of milliseconds)
        clocks = ClockSampler(local_rank)
        clocks.start()
        # ---------------- per-call breakdown (CUDA events around every call, eager launches): roofline of the dominant call
        bsteps = max(3, min(args.steps, 10))
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in devin] for _ in range(bsteps)]
        barrier()
        for s in range(bsteps):
            for i, (n, t) in enumerate(devin):
                ev[s][i][0].record()
                run_call(n, t)
                ev[s][i][1].record()
        barrier()
        per_call_ms = [sum(ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(bsteps)) / bsteps for i in range(len(devin))]

        # ---------------- device-resident throughput (value): K steps, the step's launches replayed from a CUDA graph
        # (the small pyramid levels are launch-bound; capturing them is what a serving loop would do).  Falls back to
        # eager launches if capture is not possible.
        launch_mode = "eager"
        graph = None
        if not args.no_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for n, t in devin:
                        run_call(n, t)
                torch.cuda.current_stream().wait_stream(side)
                # The 17 calls of a step are independent ops: capture them as three parallel branches so the
                # launch-latency-bound small levels overlap the HBM-bound large ones (branch 0: image splats and the two
                # large correlation levels; branch 1: feature splats; branch 2: small correlation levels).
                def branch_of(name):
                    if name == "splat_image" or name in ("corr_C32", "corr_C64"):
                        return 0
                    return 1 if name.startswith("splat") else 2
                branches = [torch.cuda.Stream() for _ in range(2)]
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    main = torch.cuda.current_stream()
                    for b in branches:
                        b.wait_stream(main)
                    graph_outs = []
                    for n, t in devin:
                        bi = branch_of(n)
                        if bi == 0:
                            graph_outs.append(run_call(n, t))
                        else:
                            with torch.cuda.stream(branches[bi - 1]):
                                graph_outs.append(run_call(n, t))
                    for b in branches:
                        main.wait_stream(b)
                graph.replay()
                torch.cuda.synchronize()
                launch_mode = "cuda_graph, 3 parallel branches (large calls | feature splats | small correlation levels)"
            except Exception as exc:   # noqa: BLE001 - keep the bench alive, say what happened
                graph = None
                launch_mode = f"eager (graph capture failed: {type(exc).__name__})"
                torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                for n, t in devin:
                    run_call(n, t)
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        # the same K steps with eager launches on one stream (what a caller that does not capture graphs gets)
        barrier()
        e0.record()
        for s in range(args.steps):
            for n, t in devin:
                run_call(n, t)
        e1.record()
        barrier()
        ms_eager_total = e0.elapsed_time(e1)

        # ---------------- end to end: pinned host -> device -> ops -> pinned host, EVERY step, through the drop-in API.
        # Three streams (H2D / compute / D2H) and two buffer sets, so step k+1's upload and step k-1's download overlap
        # step k's kernels; nothing is skipped: every step uploads all inputs and downloads all outputs.
        # The ~50 tensors of a step are staged in ONE pinned arena per direction (every tensor a 256-byte aligned view), so a step
        # is one cudaMemcpyAsync up and one down: some boxes of the pool charge ~0.4 ms per asynchronous copy, which cost the
        # tensor-by-tensor version (kept below as `e2e_many_copies`) half its throughput there and nothing elsewhere.
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_comp = torch.cuda.current_stream()

        def carve(arena, shapes):
            views, off = [], 0
            for shp in shapes:
                n_el = 1
                for d_ in shp:
                    n_el *= d_
                views.append(arena[off:off + n_el].view(shp))
                off += (n_el + 63) // 64 * 64
            return views

        def arena_size(shapes):
            tot = 0
            for shp in shapes:
                n_el = 1
                for d_ in shp:
                    n_el *= d_
                tot += (n_el + 63) // 64 * 64
            return tot

        in_keys = [(i, k) for i, (n, t) in enumerate(host) for k, v in t.items() if v is not None]
        in_shapes = [tuple(host[i][1][k].shape) for i, k in in_keys]
        pin_in = torch.empty(arena_size(in_shapes), dtype=torch.float32).pin_memory()
        for view, (i, k) in zip(carve(pin_in, in_shapes), in_keys):
            view.copy_(host[i][1][k])
        dev_in = [torch.empty(pin_in.numel(), dtype=torch.float32, device=dev) for _ in range(2)]
        arena_sets = []
        for b_ in range(2):
            views = dict(zip(in_keys, carve(dev_in[b_], in_shapes)))
            arena_sets.append([(n, {k: (None if v is None else views[(i, k)]) for k, v in t.items()}) for i, (n, t) in enumerate(host)])
        out_shapes = [tuple(o.shape) for o in [run_call(n, t) for n, t in arena_sets[0]]]      # shapes only (the arena holds garbage yet)
        dev_out = [torch.empty(arena_size(out_shapes), dtype=torch.float32, device=dev) for _ in range(2)]
        pin_out = [torch.empty(dev_out[0].numel(), dtype=torch.float32).pin_memory() for _ in range(2)]
        dev_out_views = [carve(dev_out[b_], out_shapes) for b_ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        used = [False, False]

        def e2e_step(k):
            b = k & 1
            with torch.cuda.stream(s_in):
                if used[b]:
                    s_in.wait_event(ev_comp[b])            # the kernels of step k-2 have finished reading this input set
                dev_in[b].copy_(pin_in, non_blocking=True)
                ev_in[b].record(s_in)
            s_comp.wait_event(ev_in[b])
            if used[b]:
                s_comp.wait_event(ev_out[b])               # step k-2's download has finished reading this output arena
            outs = [run_call(n, t) for n, t in arena_sets[b]]
            for view, o in zip(dev_out_views[b], outs):    # results gathered into the output arena (device-to-device, 0.78 GB)
                view.copy_(o)
            ev_comp[b].record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[b])
                pin_out[b].copy_(dev_out[b], non_blocking=True)
                ev_out[b].record(s_out)
            used[b] = True

        e2e_warm = 2
        e2e_steps = max(2, min(args.steps, 10))
        for k in range(e2e_warm):
            e2e_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_step(k)
        barrier()
        e2e_s = time.perf_counter() - t0
        d2h_bytes = sum(4 * o.numel() for o in dev_out_views[0])
        # the result that reached the host is the result the device computed (first output of the last step)
        assert torch.equal(carve(pin_out[(e2e_steps - 1) & 1], out_shapes)[0], dev_out_views[(e2e_steps - 1) & 1][0].cpu())
        del dev_in, dev_out, pin_out, arena_sets, dev_out_views

        # the same leg tensor by tensor (one asynchronous copy per tensor: ~50 up and 17 down per step), reported beside it
        dev_sets = [[(n, {k: (None if v is None else torch.empty_like(v, device=dev)) for k, v in t.items()}) for n, t in pinned]
                    for _ in range(2)]
        out_sets = [None, None]
        used = [False, False]

        def many_step(k):
            b = k & 1
            with torch.cuda.stream(s_in):
                if used[b]:
                    s_in.wait_event(ev_comp[b])
                for (n, src), (_, dst) in zip(pinned, dev_sets[b]):
                    for key, v in src.items():
                        if v is not None:
                            dst[key].copy_(v, non_blocking=True)
                ev_in[b].record(s_in)
            s_comp.wait_event(ev_in[b])
            outs = [run_call(n, t) for n, t in dev_sets[b]]
            ev_comp[b].record(s_comp)
            if out_sets[b] is None:
                out_sets[b] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[b])
                for oh, o in zip(out_sets[b], outs):
                    o.record_stream(s_out)
                    oh.copy_(o, non_blocking=True)
                ev_out[b].record(s_out)
            used[b] = True

        for k in range(e2e_warm):
            many_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            many_step(k)
        barrier()
        e2e_many_s = time.perf_counter() - t0
        del dev_sets, out_sets
        used = [False, False]

        # ---------------- e2e_frames: what a real pipeline moves over PCIe - the two frames up, the two warped frames down;
        # flows, metrics, features and cost volumes are produced and consumed on the device (fLDRnet.py:368-453).  Same
        # three-stream scheme; every step uploads both frames and downloads both splatted frames.
        img_idx = [i for i, (n, _) in enumerate(pinned) if n == "splat_image"]
        fr_dev = [[torch.empty_like(pinned[i][1]["x"], device=dev) for i in img_idx] for _ in range(2)]
        fr_out = [[torch.empty(pinned[i][1]["x"].shape, dtype=torch.float32).pin_memory() for i in img_idx] for _ in range(2)]
        fused = [False, False]

        def frames_step(k):
            b = k & 1
            with torch.cuda.stream(s_in):
                if fused[b]:
                    s_in.wait_event(ev_comp[b])
                for j, i in enumerate(img_idx):
                    fr_dev[b][j].copy_(pinned[i][1]["x"], non_blocking=True)
                ev_in[b].record(s_in)
            s_comp.wait_event(ev_in[b])
            outs = []
            j = 0
            for i, (n, t) in enumerate(devin):
                if n == "splat_image":
                    outs.append(splat(fr_dev[b][j], t["flow"], t["z"]))
                    j += 1
                else:
                    run_call(n, t)
            ev_comp[b].record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[b])
                for oh, o in zip(fr_out[b], outs):
                    o.record_stream(s_out)
                    oh.copy_(o, non_blocking=True)
                ev_out[b].record(s_out)
            fused[b] = True

        for k in range(2):
            frames_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            frames_step(k)
        barrier()
        e2e_frames_s = time.perf_counter() - t0
        frames_bytes = sum(pinned[i][1]["x"].numel() * 4 for i in img_idx)

        # measured host <-> device copy ceiling of this rank (256 MiB pinned buffer, one direction at a time and both at once)
        cbuf_h = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
        cbuf_h2 = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
        cbuf_d, cbuf_d2 = torch.empty(64 << 20, dtype=torch.float32, device=dev), torch.empty(64 << 20, dtype=torch.float32, device=dev)

        def wall(fn, reps=4):
            fn(); torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def both_ways():
            with torch.cuda.stream(s_in):
                cbuf_d.copy_(cbuf_h, non_blocking=True)
            with torch.cuda.stream(s_out):
                cbuf_h2.copy_(cbuf_d2, non_blocking=True)
        t_h2d = wall(lambda: cbuf_d.copy_(cbuf_h, non_blocking=True))
        t_d2h = wall(lambda: cbuf_h2.copy_(cbuf_d2, non_blocking=True))
        t_both = wall(both_ways)
        nb = cbuf_h.numel() * 4
        copy_ceiling = {"h2d_GBps": nb / t_h2d / 1e9, "d2h_GBps": nb / t_d2h / 1e9, "both_directions_GBps_each": nb / t_both / 1e9}
        del cbuf_h, cbuf_h2, cbuf_d, cbuf_d2

        # ---------------- strong scaling (BASELINE configs[3]): 64 DISTINCT frame pairs sharded by pair over the ranks
        # (fldr_vfi_b200.sharding.run_sharded: pair i -> rank i mod N, no collective).  Pair i's inputs are this rank's
        # tensors rolled by 8 i columns (built on the device, outside the timed events); each pair's step is timed with
        # CUDA events; a rank's time is the sum over its pairs, the job's time the max over ranks.
        from fldr_vfi_b200 import sharding
        n_pairs = 64

        def run_pair(i):
            inputs = [(n, {k: (None if v is None else torch.roll(v, shifts=8 * i, dims=-1)) for k, v in t.items()}) for n, t in devin]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for n, t in inputs:
                run_call(n, t)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        barrier()
        pair_ms = sharding.run_sharded(list(range(n_pairs)), run_pair, rank, world)
        strong_ms = sum(pair_ms.values())
        clk = clocks.stop()

    # max over ranks
    copy_min = dict(copy_ceiling)
    if world > 1:
        tt = torch.tensor([ms_total, e2e_s, ms_eager_total, e2e_frames_s, strong_ms, e2e_many_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, ms_eager_total, e2e_frames_s, strong_ms, e2e_many_s = (float(v) for v in tt)
        cc = torch.tensor([copy_ceiling["h2d_GBps"], copy_ceiling["d2h_GBps"], copy_ceiling["both_directions_GBps_each"]], device=dev, dtype=torch.float64)
        dist.all_reduce(cc, op=dist.ReduceOp.MIN)
        copy_min = {"h2d_GBps": float(cc[0]), "d2h_GBps": float(cc[1]), "both_directions_GBps_each": float(cc[2])}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_per_step = ms_total / args.steps
        # dominant call = largest share of the step; its roofline from algorithmic bytes / its own event time
        agg = {}
        for n, ms, (_, t) in zip(names, per_call_ms, host):
            a = agg.setdefault(n, [0.0, 0, 0])
            a[0] += ms
            a[1] += call_bytes(n, t)
            a[2] += 1
        dom = max(agg, key=lambda k: agg[k][0])
        dom_ms = agg[dom][0] / agg[dom][2]
        dom_bytes = agg[dom][1] / agg[dom][2]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        breakdown = {k: {"ms_per_call": round(v[0] / v[2], 4), "calls": v[2], "GBps": round(v[1] / v[2] / (v[0] / v[2] * 1e-3) / 1e9, 1),
                         "frac_of_peak": round(v[1] / v[2] / (v[0] / v[2] * 1e-3) / 1e9 / peak, 3)} for k, v in agg.items()}
        line = {
            "metric": METRIC, "value": world * 1000.0 / ms_per_step, "unit": "frame-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(),
            "notes": {"l2_policy": "inputs+outputs per step (>1.8 GB) exceed the 126 MB L2; no explicit flush",
                      "launch_mode": launch_mode,
                      "value_eager": {"value": world * 1000.0 * args.steps / ms_eager_total, "unit": "frame-pairs/s",
                                      "ms_per_step": ms_eager_total / args.steps,
                                      "what": "the same K steps with eager launches on one stream (no graph)"},
                      "e2e_mode": "H2D / compute / D2H on three streams, two buffer sets, every step copies all inputs and outputs (one pinned arena per direction, results gathered on the device); "
                                  "each rank is bound to its own slice of the cores nearest its GPU before it allocates pinned memory",
                      "host_cores_of_rank0": cores},
            "e2e": {"value": world * e2e_steps / e2e_s, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "per_rank_GBps": {"h2d": h2d_bytes * e2e_steps / e2e_s / 1e9, "d2h": d2h_bytes * e2e_steps / e2e_s / 1e9},
                    "copy_ceiling_min_over_ranks": {k: round(v, 2) for k, v in copy_min.items()},
                    "copies_per_step": {"h2d": 1, "d2h": 1},
                    "many_copies": {"value": world * e2e_steps / e2e_many_s, "unit": "frame-pairs/s",
                                    "what": "the same leg with one asynchronous copy per tensor (~50 up, 17 down per step) instead of one staged arena per direction"}},
            "e2e_frames": {"value": world * e2e_steps / e2e_frames_s, "unit": "frame-pairs/s", "h2d_bytes_per_step": frames_bytes,
                           "d2h_bytes_per_step": frames_bytes, "steps": e2e_steps,
                           "what": "same step, but only the two frames cross PCIe (up) and the two splatted frames (down); flows, "
                                   "metrics, features and cost volumes stay on the device as in fLDRnet's forward"},
            "strong_scaling": {"pairs": n_pairs, "seconds": strong_ms / 1e3, "value": n_pairs / (strong_ms / 1e3), "unit": "frame-pairs/s",
                               "pairs_per_rank": sharding.shard_counts(n_pairs, world),
                               "what": "BASELINE configs[3]: 64 distinct pairs, pair i -> rank i mod N (sharding.run_sharded), device-timed "
                                       "per pair (eager launches), rank time = sum over its pairs, job time = max over ranks"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": ncu_traffic(dom), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(dom_bytes), "ms_per_launch": round(dom_ms, 4)},
            "breakdown": breakdown,
            "clocks": clk,
        }
        if world == 1:
            line["flow_regimes"] = flow_regimes_probe(splat, devin, peak)
            line["train_step"] = train_step_probe(peak)
            line["next_rows"] = next_rows_probe(devin, peak)
        if world == 1 and not args.no_fldrnet:
            line["fldrnet_e2e"] = fldrnet_e2e_probe()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(frac=1.0)
        if world == 1 and not args.no_ref_gpu:
            line["ref_gpu"] = ref_gpu_baseline(devin, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- further legs (rank 0, N = 1)
def flow_regimes_probe(splat, devin, peak):
    """The 4K image splat under the three flow regimes of SURVEY.md 8d (the headline step uses F1): F1 smooth, F2 iid
    U(-64, 64) px per pixel (no merging of reductions possible), F3 every pixel converges on the frame centre (maximum
    contention).  Device-resident, CUDA events, median of 20; inputs exceed the L2."""
    from oracle import synth   # input generator only
    t = [t for n, t in devin if n == "splat_image"][0]
    N, C, H, W = t["x"].shape
    nbytes = splat_bytes(N, C, H, W, True)
    out = {}
    with torch.no_grad():
        for reg, seed in (("F1", 57), ("F2", 59), ("F3", 60)):
            fl = synth.flow(N, H, W, reg, seed=seed).to(t["x"].device)
            ms, lo, hi = median_event_ms(lambda: splat(t["x"], fl, t["z"]))
            out[reg] = {"ms_per_call": round(ms, 4), "min_ms": round(lo, 4), "max_ms": round(hi, 4),
                        "GBps": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
        # two more points on the same axis (not SURVEY regimes): zero flow, and F1's amplitude (sigma 16 px) on a control grid 8x
        # coarser (local stretch ~4 % instead of ~35 %: closer to real inter-frame motion).  The scatter pass is bound by the
        # number of 32-byte sectors its reductions touch, which is what the flow's local stretch sets.
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(61)
        lo_res = torch.randn(N, 2, max(H // 512, 2), max(W // 512, 2), generator=g) * (16.0 * W / 4096.0)
        extra = {"F0_zero": torch.zeros(N, 2, H, W), "F1_gentle_stretch": F.interpolate(lo_res, size=(H, W), mode="bilinear", align_corners=False).contiguous()}
        for reg, fl_host in extra.items():
            fl = fl_host.to(t["x"].device)
            ms, lo, hi = median_event_ms(lambda: splat(t["x"], fl, t["z"]))
            out[reg] = {"ms_per_call": round(ms, 4), "min_ms": round(lo, 4), "max_ms": round(hi, 4),
                        "GBps": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    return out


def train_step_probe(peak):
    """BASELINE configs[4]: the train_it.py-shaped step (main.py:545-660, utils.py:848-864) at its literal shapes - 32 synthetic
    512x512 crops: image splat 32x3x512^2 with metric (input / flow / metric gradients), feature splat 32x48x64^2 (input
    gradient; the flow is detached, fLDRnet.py:384), correlation B = 64 at (32,128,128) and (64,64,64) (both gradients).
    Forward and backward are timed separately (the backward through autograd on a retained graph), CUDA events, median of
    10; every kernel carries its roofline entry: algorithmic bytes of SURVEY.md 8d / time / measured HBM peak."""
    import fldr_vfi_b200.correlation as C
    import fldr_vfi_b200.softSplat as S
    from oracle import synth   # input generator only
    dev = torch.device("cuda", torch.cuda.current_device())
    sp = S.Softsplat()
    out = {}

    def entry(ms, nbytes, what):
        return {"ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GBps": round(nbytes / ms / 1e6, 1),
                "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3), "bytes_formula": what}

    def splat_case(tag, N, Cc, H, W, metric, need_flow, scale):
        x = (synth.image(N, Cc, H, W, seed=1) if Cc == 3 else synth.features(N, Cc, H, W, seed=1)).to(dev).requires_grad_(True)
        fl = (synth.flow(N, H, W, "F1", seed=2) * scale).to(dev).requires_grad_(need_flow)
        z = synth.metric(N, H, W, seed=3).to(dev).requires_grad_(True) if metric else None
        g = synth.grad((N, Cc, H, W), seed=4).to(dev)
        with torch.no_grad():
            fwd, _, _ = median_event_ms(lambda: sp(x, fl, z), iters=10)
        y = sp(x, fl, z)
        wrt = [t for t in (x, fl, z) if t is not None and t.requires_grad]
        bwd_autograd, _, _ = median_event_ms(lambda: torch.autograd.grad(y, wrt, g, retain_graph=True), iters=10)
        # the backward is two short kernels: through torch.autograd.grad the timed region is bound by the autograd engine's host
        # time, so the op-level entry the autograd Function calls (softSplat._splat_backward) is what the roofline entry times
        with torch.no_grad():
            yy, norm = S._splat_forward(3, x, fl, z, True)
            need = (True, need_flow, metric)
            bwd, _, _ = median_event_ms(lambda: S._splat_backward(3, x, fl, z, yy, norm, g, need), iters=10)
        px = N * H * W
        m = 1 if metric else 0
        out[tag + "_fwd"] = entry(fwd, 4 * px * (2 * Cc + 2 + m), "4*NHW*(2C+2+[metric])")
        if need_flow:
            out[tag + "_bwd"] = entry(bwd, 4 * px * (4 * Cc + 7), "4*NHW*(4C+7): all three gradients")
        else:
            out[tag + "_bwd"] = entry(bwd, 4 * px * (3 * Cc + 3), "4*NHW*(3C+3): grad_input only (reads gOut, out, flow, norm)")
        out[tag + "_bwd"]["ms_through_autograd_grad"] = round(bwd_autograd, 4)

    def corr_case(tag, B, Cc, H, W):
        a = synth.features(B, Cc, H, W, seed=3).to(dev).requires_grad_(True)
        b = synth.features(B, Cc, H, W, seed=5).to(dev).requires_grad_(True)
        g = synth.grad((B, 81, H, W), seed=4).to(dev)
        with torch.no_grad():
            fwd, _, _ = median_event_ms(lambda: C.FunctionCorrelation(tensorFirst=a, tensorSecond=b), iters=10)
        o = C.FunctionCorrelation(tensorFirst=a, tensorSecond=b)
        bwd, _, _ = median_event_ms(lambda: torch.autograd.grad(o, [a, b], g, retain_graph=True), iters=10)
        px = B * H * W
        out[tag + "_fwd"] = entry(fwd, 4 * px * (2 * Cc + 81), "4*BHW*(2C+81)")
        out[tag + "_fwd"]["TFLOPs"] = round(162 * Cc * px / fwd / 1e9, 2)
        out[tag + "_bwd"] = entry(bwd, 4 * px * (4 * Cc + 81), "4*BHW*(4C+81): both gradients")
        out[tag + "_bwd"]["TFLOPs"] = round(324 * Cc * px / bwd / 1e9, 2)

    splat_case("splat_image_32x3x512x512_metric", 32, 3, 512, 512, True, True, 4.0)
    splat_case("splat_feat_32x48x64x64", 32, 48, 64, 64, False, False, 16.0)
    corr_case("corr_64x32x128x128", 64, 32, 128, 128)
    corr_case("corr_64x64x64x64", 64, 64, 64, 64)
    fwd_bwd = sum(v["ms"] for v in out.values())
    return {"kernels": out, "ms_sum": round(fwd_bwd, 3),
            "what": "device-resident, inputs of each call exceed or fill the L2; correlation backward timed through autograd on a retained graph, splat backward at the op-level entry the autograd Function calls (its time through torch.autograd.grad beside it)"}


def fldrnet_e2e_probe():
    """BASELINE configs[2] / north_star target 2: the UNTOUCHED fLDRnet (--papermodel --test5scales, shipped checkpoint) on one
    synthetic 4096x2160 triplet, model forward wall time (synchronised both sides), 5 repetitions after one warm-up, each
    variant in its own process: (reference) the reference's CuPy kernels through the NVRTC shim, (ours) softSplat +
    correlation replaced by the drop-ins, (ours_warp) + the bwarp method of SURVEY 8f-1 swapped on the imported class,
    (ours_rows) + the to_pca_diff name of SURVEY 8f-4 swapped in the imported module.
    PSNR against the synthetic ground truth per variant."""
    import subprocess
    script = os.path.join(ROOT, "baseline", "e2e_fldrnet.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "checkpoint_dir")):
        return {"unavailable": "baseline/_ref not staged (python baseline/fetch_ref.py in the build container)"}
    try:
        r = subprocess.run([sys.executable, script, "--reps", "5", "--variants", "reference,ours,ours_warp,ours_rows"], capture_output=True,
                           text=True, timeout=900)
        last = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
        if r.returncode != 0 or not last:
            return {"unavailable": ("rc %d: " % r.returncode) + (r.stderr or r.stdout)[-300:]}
        return json.loads(last[-1])
    except Exception as exc:      # a reported leg must never take the bench down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


# ----------------------------------------------------------------------------------------------- reference on the GPU
def next_rows_probe(devin, peak):
    """SURVEY.md 8f rank 1 (outside the step above, reported beside it): the backward warp and the splat metric that
    fLDRnet runs just before the image splat (fLDRnet.py:442-446, 546-581), on this pair's 4K images and flows.
    Inputs exceed the L2; CUDA events on the current stream; median of 20."""
    import fldr_vfi_b200.warp as Wp
    img = [t for n, t in devin if n == "splat_image"]
    x0, x1, fl = img[0]["x"], img[1]["x"], img[0]["flow"]
    N, C, H, W = x0.shape
    px = N * H * W

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    out = {}
    with torch.no_grad():
        for name, nbytes, fn in (("bwarp_image_C3", 4 * px * (C + 2 + C), lambda: Wp.bwarp(x1, fl, True)),
                                 ("splat_metric_C3", 4 * px * (2 * C + 2 + 1), lambda: Wp.splat_metric(x0, x1, fl, -1.894))):
            ms = med(fn)
            out[name] = {"ms_per_call": round(ms, 4), "algorithmic_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1),
                         "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    # rank 2: occlusion softmax + six-way blend (fLDRnet.py:510-524), float64 output
    import fldr_vfi_b200.blend as Bl
    from oracle import synth   # input generator only
    logits = (synth.grad((N, 6, H, W), seed=77) * 3.0).to(x0.device)
    extra = [synth.image(N, C, H, W, seed=78 + k).to(x0.device) for k in range(4)]
    tv = torch.full((N, 1, 1, 1), 0.5, device=x0.device)
    T = torch.ones(1, dtype=torch.float64, device=x0.device)
    with torch.no_grad():
        ms = med(lambda: Bl.occ_blend(logits, T, tv, *extra, x0, x1))
    nbytes = px * (6 * 4 + 6 * C * 4 + C * 8)
    out["occ_blend_C3_f64"] = {"ms_per_call": round(ms, 4), "algorithmic_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1),
                               "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    # rank 4: block-PCA features of the two stacked frames (pca_comp.py:473-528), float32 result (the .float() of fLDRnet.py:146)
    import fldr_vfi_b200.pca as Pc
    g = torch.Generator().manual_seed(5)
    mean = (torch.randn(64, generator=g, dtype=torch.float64) * 0.1).to(x0.device)
    EV = torch.linalg.qr(torch.randn(64, 64, generator=g, dtype=torch.float64))[0][:16].contiguous().to(x0.device)
    mv = (torch.rand(16, generator=g, dtype=torch.float64) + 0.5).to(x0.device)
    im6 = torch.cat([x0[0], x1[0]], 0)
    with torch.no_grad():
        ms = med(lambda: Pc.pca_features(im6, mean, EV, mv, out_dtype=torch.float32))
    nbytes = 6 * H * W * 4 + 6 * 16 * (H // 8) * (W // 8) * 4
    out["pca_features_6xHxW_f32out"] = {"ms_per_call": round(ms, 4), "algorithmic_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1),
                                        "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3),
                                        "bound_note": "1.8 GFLOP of float64 per call on the float64 tensor cores (mma.sync.m8n8k4): 49 us at the measured 37 TFLOP/s, 43 us at the HBM roofline; two launches (projection + min/max, rescale)"}
    try:      # the reference's own function text on the same GPU (torch operator sequence incl. a float64 matmul)
        import types
        from baseline import ref_src
        if ref_src.available():
            ref_fn = ref_src.to_pca_diff()
            prm = types.SimpleNamespace(wiS=8, weightMat=None, components_fraction=0.25)
            rargs = types.SimpleNamespace(gpu=x0.device, mean_vector_norm=True)
            with torch.no_grad():
                ms_ref = med(lambda: ref_fn(im6, prm, rargs, mean, EV, mv).float())
            out["pca_features_6xHxW_f32out"]["reference_text_same_gpu_ms"] = round(ms_ref, 4)
    except Exception as exc:      # a reported baseline must never take the bench down
        out["pca_features_6xHxW_f32out"]["reference_text_same_gpu_ms"] = f"unavailable: {type(exc).__name__}"
    # rank 4, second half: the input pyramid (main.py:855-856) - five bicubic levels of the two padded frames in one launch
    import fldr_vfi_b200.pyramid as Py
    import torch.nn.functional as F
    frames5 = torch.stack([x0, x1], 2).contiguous()                   # [N, C, T=2, H, W]
    scales = [8, 16, 32, 64, 128, 256]
    def med_dev(fn, reps=4):     # calls shorter than their host-side launch cost: let the host run ahead behind a spin kernel
        fn()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        return sorted(ts)[len(ts) // 2]

    with torch.no_grad():
        ms = med_dev(lambda: Py.input_pyramid(frames5, scales, 5))
    nbytes = int(4 * N * C * 2 * sum((H >> k) * (W >> k) for k in range(6)))
    out["input_pyramid_5_levels"] = {"ms_per_call": round(ms, 4), "algorithmic_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1),
                                     "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    try:
        planes = frames5.permute(0, 2, 1, 3, 4).reshape(N * 2, C, H, W)
        with torch.no_grad():       # torch's own CUDA bicubic, one call per level (what moving the reference's expression to the GPU gives)
            out["input_pyramid_5_levels"]["torch_cuda_interpolate_ms"] = round(
                med_dev(lambda: [F.interpolate(planes, scale_factor=1.0 / (1 << k), mode="bicubic", align_corners=False) for k in range(1, 6)]), 4)
        host = planes.cpu()
        t0 = time.perf_counter()   # the reference as written: CPU bicubic per level + a pageable H2D copy per level (one repetition)
        with torch.no_grad():
            lv = [F.interpolate(host, scale_factor=1.0 / (1 << k), mode="bicubic", align_corners=False).to(x0.device) for k in range(1, 6)]
        torch.cuda.synchronize()
        out["input_pyramid_5_levels"]["reference_expression_cpu_plus_h2d_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        del lv
    except Exception as exc:
        out["input_pyramid_5_levels"]["torch_cuda_interpolate_ms"] = f"unavailable: {type(exc).__name__}"
    return out


def ref_gpu_baseline(devin, args):
    """The reference's own CuPy kernels on the same B200 (north_star's first reported baseline): the UNMODIFIED
    softSplat.py / OpticalFlow/correlation.py staged in baseline/_ref, CuPy replaced by an NVRTC shim
    (baseline/cupy_shim).  Same inputs, same step, device-resident, CUDA events.  A reported baseline, not a target."""
    try:
        from baseline import ref_gpu
        if not ref_gpu.available():
            return {"unavailable": "baseline/_ref not staged (python baseline/fetch_ref.py in the build container)"}
        RS = ref_gpu.softsplat_module()
        RC = ref_gpu.correlation_module()
        splat = RS.Softsplat()

        def run_call(name, t):
            if name.startswith("splat"):
                return splat(t["x"], t["flow"], t["z"])
            return RC.FunctionCorrelation(tensorFirst=t["f1"], tensorSecond=t["f2"])

        steps = max(1, min(args.steps, 5))
        with torch.no_grad():
            for _ in range(3):                      # NVRTC compiles per shape on first use
                for n, t in devin:
                    run_call(n, t)
            ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in devin] for _ in range(steps)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for s in range(steps):
                for i, (n, t) in enumerate(devin):
                    ev[s][i][0].record()
                    run_call(n, t)
                    ev[s][i][1].record()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        agg = {}
        for i, (n, _) in enumerate(devin):
            a = agg.setdefault(n, [0.0, 0])
            a[0] += sum(ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(steps)) / steps
            a[1] += 1
        return {"value": 1000.0 / ms, "unit": "frame-pairs/s", "ms_per_step": ms, "steps": steps,
                "what": "unmodified reference op files + their CuPy kernel strings via NVRTC on this GPU, whole "
                        "FunctionSoftsplat / FunctionCorrelation calls (host templating included)",
                "ms_per_call": {k: round(v[0] / v[1], 4) for k, v in agg.items()}}
    except Exception as exc:          # a reported baseline must never take the bench down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_step(calls):
    """The reference's own arithmetic on host cores: splat through the reference kernel text compiled for the
    host (oracle/_ref) inside the restated torch glue; correlation through the torch restatement (the
    reference's updateOutput kernel needs block barriers - its fiber emulation is far too slow to time)."""
    from oracle import corr_oracle, ref_host, splat_oracle
    use_ref = ref_host.available()
    outs = []
    for n, t in calls:
        if n.startswith("splat"):
            f = ref_host.function_softsplat if use_ref else splat_oracle.function_softsplat
            outs.append(f(t["x"], t["flow"], t["z"], "softmax"))
        else:
            outs.append(corr_oracle.correlation_fwd(t["f1"], t["f2"]))
    return outs, use_ref


def cpu_baseline(frac):
    torch.set_num_threads(os.cpu_count() or 1)
    calls = make_pair_inputs(seed=0, frac=frac)
    with torch.no_grad():
        cpu_step(calls)
        t0 = time.perf_counter()
        _, use_ref = cpu_step(calls)
        dt = time.perf_counter() - t0
    return {"value": frac / dt, "unit": "frame-pairs/s", "cores": os.cpu_count(),
            "kind": "reference" if use_ref else "port",
            "kind_detail": "reference (splat) + port (correlation)" if use_ref else "port",
            "sample": f"top {frac:.3f} of the rows of every tensor of one 4K frame pair (same 17 calls), {dt:.2f} s; "
                      "splat = reference kernel text on host cores (oracle/_ref, OpenMP) + restated torch glue, "
                      "correlation = torch restatement of the reference formulas"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    # bounded sample: choose the row fraction so (steps + warmup) stays within ~2.5 minutes
    frac = 1.0
    calls = make_pair_inputs(seed=0, frac=frac)
    with torch.no_grad():
        t0 = time.perf_counter()
        cpu_step(calls)
        t_one = time.perf_counter() - t0
        budget = 150.0
        while frac > 1.0 / 64 and t_one * (args.steps + args.warmup) > budget:
            frac /= 2
            calls = make_pair_inputs(seed=0, frac=frac)
            t0 = time.perf_counter()
            cpu_step(calls)
            t_one = time.perf_counter() - t0
        for _ in range(args.warmup):
            cpu_step(calls)
        t0 = time.perf_counter()
        use_ref = False
        for _ in range(args.steps):
            _, use_ref = cpu_step(calls)
        dt = time.perf_counter() - t0
    value = frac * args.steps / dt
    sample = (f"top {frac:.4f} of the rows of every tensor of one 4K frame pair per step (same 17 calls); splat = reference "
              "kernel text on host cores (oracle/_ref, OpenMP) + restated torch glue, correlation = torch restatement")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(),
            "notes": {"reference_arm": "reference (splat: the reference's kernel text on host cores) + port (correlation: torch "
                                       "restatement of the reference formulas - its updateOutput kernel needs block barriers)"},
            "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": os.cpu_count(),
                             "kind": "reference" if use_ref else "port",
                             "kind_detail": "reference (splat) + port (correlation)" if use_ref else "port", "sample": sample},
            "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-fldrnet", action="store_true", help="skip the untouched-fLDRnet end-to-end leg (about two minutes)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay of the step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

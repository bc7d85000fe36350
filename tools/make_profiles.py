"""Turn what tools/gpu_r2_profiles.sh brought back in gpurun_out/ into the tracked round-2 summaries under profiles/ (runs on the
CPU box: ncu -i reads the reports, cuobjdump reads the shipped library).  python tools/make_profiles.py"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarise_profiles  # noqa: E402

KEEP = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]


def first_of_each(rows):
    seen, out = set(), []
    for r in rows:
        import math
        key = (r["Kernel Name"], r["Grid Size"], round(math.log2(max(float(r["gpu__time_duration.sum"]), 1.0)) * 2))   # same kernel and grid on another level = another row
        if key not in seen:
            seen.add(key)
            out.append(r)
    return out


def main():
    sections = [("prof_r2_corr.ncu-rep", "python bench.py --steps 1 --warmup 0: correlation forward, the five levels of the native-4K pyramid"),
                ("prof_r2_corrbwd.ncu-rep", "python tools/corr_bwd_prof.py: correlation backward at 64x32x128x128, one persistent launch per gradient (gradFirst, gradSecond)"),
                ("prof_r2_top.ncu-rep", "python bench.py --steps 1 --warmup 0 (cold caches, serialised launches): first launch of every (kernel, grid) of one 4K frame-pair step"),
                ("prof_r2_train.ncu-rep", "python tools/train_probe.py (cfg5 training shapes): splat backward (gS prep + corner gather) and correlation backward"),
                ("prof_r2_next.ncu-rep", "python tools/next_rows_ncu.py: the next-row kernels (SURVEY 8f) at the padded 4K frame pair")]
    traffic = {"_comment": "DRAM traffic per call from ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum of the call's launches, cold caches), round 2 final; "
                           "read by bench.py for roofline.traffic.  Source: profiles/r2_ncu_full_top_kernels.txt.  Dirty accumulator lines still in L2 when a kernel ends are "
                           "written back later and are not in these sums."}
    with open(os.path.join(PROF, "r2_ncu_full_top_kernels.txt"), "w") as f:
        for rep, title in sections:
            path = os.path.join(OUT, rep)
            if not os.path.exists(path):
                continue
            rows = first_of_each(raw_rows(path))
            f.write(f"# ncu --set full --clock-control none, {title}\n")
            for r in rows:
                f.write(json.dumps({k: r[k] for k in KEEP if k in r}) + "\n")
            def tot(pred):
                sel = [r for r in rows if pred(r)]
                return sel, sum(float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"]) for r in sel) * 1e6
            if rep == "prof_r2_corr.ncu-rep":
                sel, b = tot(lambda r: "corr81_fwd_tma" in r["Kernel Name"] and r["Grid Size"] == "(296, 1, 1)" and float(r["dram__bytes_write.sum"]) > 200)
                if sel:
                    traffic["corr_C32"] = {"bytes": int(b / len(sel)), "kernels": {"corr81_fwd_tma_kernel": int(b / len(sel))}}
            if rep == "prof_r2_corrbwd.ncu-rep":
                sel, b = tot(lambda r: "corr81_bwd_rows" in r["Kernel Name"])
                if sel:
                    traffic["corr_bwd_64x32x128x128"] = {"bytes": int(b), "kernels": {r["Kernel Name"][:48]: int((float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6) for r in sel}}
            if rep == "prof_r2_top.ncu-rep":
                parts = {}
                for name, pat, grid in (("splat_zero_kernel", "splat_zero", "(4611, 1, 1)"), ("splat_scatter_tile_kernel", "splat_scatter_tile", "(32, 288, 1)"),
                                        ("splat_normalise_kernel", "splat_normalise", "(8, 2304, 1)")):
                    sel, b = tot(lambda r: pat in r["Kernel Name"] and (grid is None or r["Grid Size"] == grid))
                    if sel:
                        parts[name] = int(b / len(sel))
                traffic["splat_image"] = {"bytes": sum(parts.values()), "kernels": parts}
                sel, b = tot(lambda r: "corr81_fwd_tma" in r["Kernel Name"] and r["Grid Size"] == "(296, 1, 1)" and float(r["dram__bytes_write.sum"]) > 200)
                if sel:
                    traffic["corr_C32"] = {"bytes": int(b / len(sel)), "kernels": {"corr81_fwd_tma_kernel": int(b / len(sel))}}
            if rep == "prof_r2_train.ncu-rep":
                sel, b = tot(lambda r: "corr81_bwd_rows" in r["Kernel Name"])
                if sel:
                    traffic["corr_bwd_64x32x128x128"] = {"bytes": int(b), "kernels": {r["Kernel Name"][:40]: int((float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6) for r in sel}}
                sel, b = tot(lambda r: "splat_bwd" in r["Kernel Name"])
                if sel:
                    traffic["splat_bwd_32x3x512x512"] = {"bytes": int(b), "kernels": {r["Kernel Name"][:40]: int((float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6) for r in sel}}
    # the image splat's three passes hand accumulator lines to each other through L2 (alternating row order), which a capture that
    # flushes the caches between kernels cannot see: take its traffic from the --cache-control none pass (steady state, warm-up steps first)
    wcsv = os.path.join(OUT, "r2_traffic_warm.csv")
    if os.path.exists(wcsv):
        rows = [r for r in csv.reader(open(wcsv)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, gi, mi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Metric Name", "ID"))
        d = collections.OrderedDict()
        for r in rows[1:]:
            e = d.setdefault(r[ii], {"k": r[ki], "g": r[gi]})
            try:
                e[r[mi]] = float(r[vi].replace(",", ""))
            except ValueError:
                pass
        parts, times = {}, {}
        for name, pat, grid in (("splat_zero_kernel", "splat_zero", "(4611, 1, 1)"), ("splat_scatter_tile_kernel", "splat_scatter_tile", "(32, 288, 1)"),
                                ("splat_normalise_kernel", "splat_normalise", "(8, 2304, 1)")):
            sel = [e for e in d.values() if pat in e["k"] and e["g"] == grid][2:]          # skip the first call (cold)
            if sel:
                parts[name] = int(sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in sel) / len(sel))
                times[name] = round(sum(e.get("gpu__time_duration.sum", 0) for e in sel) / len(sel) / 1e3, 1)
        if len(parts) == 3:
            traffic["splat_image"] = {"bytes": sum(parts.values()), "kernels": parts, "us_under_ncu": times,
                                      "capture": "ncu --cache-control none (L2 state as in the running step), mean over the image-splat calls after the first"}
    json.dump(traffic, open(os.path.join(PROF, "r2_traffic.json"), "w"), indent=1)
    # launch list of one step
    lcsv = os.path.join(OUT, "r2_launches_step.csv")
    if os.path.exists(lcsv):
        summarise_profiles.launches(lcsv, os.path.join(PROF, "r2_launches_bench_step.txt"),
                                    "every launch of one bench step (python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ref-gpu --no-fldrnet): 2 image splats, "
                                    "10 feature splats, 5 correlation levels (the capture window also holds the start of the next phases)")
    # SASS evidence from the shipped library
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "fldr-vfi_b200", "libfldr_b200.so")], capture_output=True, text=True).stdout
    want = ["UTMALDG", "UBLKCP", "SYNCS", "REDG.E.ADD.F32x4", "RED.E.ADD", "DMMA", "DFMA", "FFMA", "LDS.128", "SHFL", "CCTL", "UTMAPF", "STG.E.128", "LDG.E.128"]
    with open(os.path.join(PROF, "r2_sass_excerpts.txt"), "w") as f:
        f.write("# cuobjdump -sass fldr-vfi_b200/libfldr_b200.so (sm_100a), round 2 final: per kernel the instruction count and the count of the mnemonics that\n"
                "# prove the Blackwell-native features (UTMALDG / UBLKCP = TMA, SYNCS = mbarrier, REDG.E.ADD.F32x4 = 128-bit reduction, DMMA = float64 tensor core)\n"
                "# plus the first occurrence of each.  No UTC*MMA / LDTM: the only dense contraction on the path is the float64 block-PCA projection (DMMA.8x8x4).\n\n")
        for blk in re.split(r"\n\s*Function : ", sass)[1:]:
            name = blk.split("\n", 1)[0].strip()
            lines = [l for l in blk.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
            ops = [re.sub(r"^.*?/\*[0-9a-f]+\*/\s+", "", l).split(";")[0].strip() for l in lines]
            cnt = collections.OrderedDict((w, sum(1 for o in ops if w in o)) for w in want)
            f.write(f"{name}\n  instructions {len(ops)}  " + "  ".join(f"{k} {v}" for k, v in cnt.items() if v) + "\n")
            for w in ("UTMALDG", "SYNCS", "REDG.E.ADD.F32x4", "DMMA"):
                hit = next((o for o in ops if w in o), None)
                if hit:
                    f.write(f"    {hit}\n")
            f.write("\n")
    print("wrote profiles/r2_ncu_full_top_kernels.txt, r2_traffic.json, r2_launches_bench_step.txt, r2_sass_excerpts.txt")


if __name__ == "__main__":
    main()

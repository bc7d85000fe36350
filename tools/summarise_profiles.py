"""Turn ncu CSV logs brought back in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import sys


def launches(csv_path, out_path, title):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, mi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Metric Name", "ID"))
    d = collections.OrderedDict()
    for r in rows[1:]:
        e = d.setdefault(r[ii], {"k": r[ki], "g": r[gi]})
        try:
            e[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    tot = sum(e.get("gpu__time_duration.sum", 0) for e in d.values())
    with open(out_path, "w") as f:
        f.write(f"# {title}\n# ncu --clock-control none; per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write(f"# {'us':>9s} {'share':>6s} {'Minst':>8s} {'dramR_MB':>9s} {'dramW_MB':>9s} {'grid':>16s}  kernel\n")
        for e in d.values():
            t = e.get("gpu__time_duration.sum", 0)
            f.write(f"{t / 1e3:11.1f} {100 * t / tot:5.1f}% {e.get('smsp__inst_executed.sum', 0) / 1e6:8.2f} "
                    f"{e.get('dram__bytes_read.sum', 0) / 1e6:9.1f} {e.get('dram__bytes_write.sum', 0) / 1e6:9.1f} {e['g']:>16s}  {e['k'][:80]}\n")
        f.write(f"# total {tot / 1e3:.1f} us over {len(d)} launches\n")


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu launch list")

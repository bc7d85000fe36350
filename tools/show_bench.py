import json, sys
d = json.loads(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.txt").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
for k, v in d["breakdown"].items():
    print(f"  {k:16s} {v['ms_per_call'] * 1000:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_peak']:.3f}")

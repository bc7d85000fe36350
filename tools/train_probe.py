"""cfg5 (train_it.py-shaped step) kernels once each, for ncu: python tools/train_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
r = bench.train_step_probe(bench.measured_peak_gbs()[0])
for k, v in r["kernels"].items():
    print(f"{k:44s} {v['ms']*1000:9.1f} us  {v['GBps']:8.1f} GB/s  frac {v['frac_of_peak']:.3f}")

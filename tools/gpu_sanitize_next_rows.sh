#!/bin/bash
# compute-sanitizer memcheck over the small-shape GPU tests of the "next" rows (warp fwd/bwd, PWC Backward, blend,
# fused correlation activation) and the correlation forward with the new epilogue variants
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_next.log \
    python -m pytest tests/test_gpu_warp.py tests/test_gpu_blend.py tests/test_gpu_corr.py -m gpu -x -q -W ignore \
    -k "not 4k and not native and not full_size" 2>&1 | tail -3
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|misaligned" gpurun_out/memcheck_next.log | head -5

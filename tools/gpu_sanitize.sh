#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck (global/shared OOB, misaligned) and racecheck (shared memory)
mkdir -p gpurun_out
SEL='golden or seeded or strided or nonfinite or needs_input or streaming_kernel_small or test_vs_oracle_seeded'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_gpu_splat.py tests/test_gpu_corr.py -m gpu -x -q -k "$SEL" 2>&1 | tail -3
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|misaligned" gpurun_out/memcheck.log | head -5
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/racecheck.log \
    python -m pytest tests/test_gpu_corr.py -m gpu -x -q -k "golden or test_vs_oracle_seeded" 2>&1 | tail -3
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | head -5

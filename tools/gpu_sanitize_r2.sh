#!/bin/bash
# Round-2 kernels under compute-sanitizer: memcheck over the small-shape tests of the splat forward / backward (tile scatter with the
# cross merges, zero-fill kernel, two-pass backward), the correlation backward (TMA rings), the block-PCA projection (DMMA) and the
# input pyramid; racecheck (shared memory) over the PCA and correlation-backward tests.
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_r2.log \
    python -m pytest tests/test_gpu_splat.py tests/test_gpu_corr.py tests/test_gpu_pca.py tests/test_gpu_pyramid.py -m gpu -x -q -W ignore \
    -k "not 4k and not native and not full_size and not 2304 and not literal and not cfg5 and not batched" 2>&1 | tail -3
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|misaligned" gpurun_out/memcheck_r2.log | head -5
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/racecheck_r2.log \
    python -m pytest tests/test_gpu_pca.py tests/test_gpu_corr.py -m gpu -x -q -W ignore -k "golden or strided or test_vs_oracle_seeded" 2>&1 | tail -3
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck_r2.log | head -5

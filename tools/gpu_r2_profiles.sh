#!/bin/bash
# Round-2 evidence: bench line with clocks, ncu launch list of one step, ncu --set full of the top kernels (inference step and
# training-shape kernels), host overhead.  Everything lands in gpurun_out/ and is summarised into profiles/ by the builder.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ref-gpu --no-fldrnet > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"splat_zero|splat_scatter_tile|splat_normalise|corr81_fwd_tma" -c 14 -o gpurun_out/prof_r2_top -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ref-gpu --no-fldrnet > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"splat_bwd|corr81_bwd" -c 8 -o gpurun_out/prof_r2_train -f python tools/train_probe.py > gpurun_out/r2_train_probe_under_ncu.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwarp_kernel|occ_blend|pca_|pyramid_" -s 6 -c 6 -o gpurun_out/prof_r2_next -f python tools/next_rows_ncu.py > /dev/null 2>&1
timeout 300 python tools/train_probe.py > gpurun_out/r2_train_probe.txt 2>&1
timeout 300 python tools/overhead_probe.py > gpurun_out/r2_host_overhead.txt 2>&1
FLDR_B200_NO_EXT=1 timeout 300 python tools/overhead_probe.py > gpurun_out/r2_host_overhead_ctypes.txt 2>&1
timeout 300 python tools/splat_probe.py > gpurun_out/r2_splat_probe.txt 2>&1
timeout 300 python tools/bwd_probe.py > gpurun_out/r2_bwd_probe.txt 2>&1
tail -c 300 gpurun_out/r2_bench_full.json; echo; cat gpurun_out/r2_train_probe.txt
# the image splat's traffic as the running step sees it (the three passes hand accumulator lines to each other through L2: no cache flush between kernels)
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"splat_zero|splat_scatter_tile|splat_normalise" -c 90 --csv --log-file gpurun_out/r2_traffic_warm.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-ref-gpu --no-fldrnet > /dev/null 2>&1

#!/bin/bash
# Round-end evidence: bench line, ncu launch list of one step, ncu --set full of the top kernels, reference-on-GPU
# baseline, e2e fLDRnet, training-shape table.  Everything lands in gpurun_out/ and is summarised into profiles/ here.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"splat_scatter_merged|splat_normalise|corr81_fwd_tma" -c 14 -o gpurun_out/prof_r1_top -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
timeout 900 python baseline/e2e_fldrnet.py --reps 3 > gpurun_out/e2e_fldrnet.json 2> gpurun_out/e2e.err
timeout 600 python tools/bench_bwd.py > gpurun_out/bwd_training_shapes.txt 2>&1
timeout 300 python tools/overhead_probe.py > gpurun_out/host_overhead.txt 2>&1
timeout 200 python tools/warp_probe.py > gpurun_out/warp_probe.txt 2>&1
timeout 200 python tools/blend_probe.py > gpurun_out/blend_probe.txt 2>&1
timeout 200 python tools/corr_act_probe.py > gpurun_out/corr_act_probe.txt 2>&1
tail -c 600 gpurun_out/bench_full.json; echo; tail -c 400 gpurun_out/e2e_fldrnet.json

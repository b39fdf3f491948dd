#!/bin/bash
# Round-2 profile captures (run under gpurun, one GPU).  Outputs under gpurun_out/.
# 1. hardware-counter sections (no source patching) on the bench's C4 frame: one frame's primary + 4 shadow launches
# 2. full set + source on a smaller frame of the same scene family
# 3. launch list of the bench command
set -x
K='regex:^primary_kernel|^shadow_light'
SECS="--section LaunchStats --section Occupancy --section SpeedOfLight --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis"
timeout 600 ncu $SECS --metrics sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k "$K" -s 10 -c 5 -f -o gpurun_out/r02_span_c4 python tools/probe.py nopeak 1000000,3840,2160,4 > gpurun_out/r02_span_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 10 -c 5 -f -o gpurun_out/r02_span_200k python tools/probe.py nopeak 200000,1920,1080,4 > gpurun_out/r02_span_ncu_200k.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_span_launches.csv python bench.py --steps 2 --warmup 1 --no-cull --no-cpu-baseline --no-e2e > gpurun_out/r02_span_ncu_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# final round-2 measurements on one GPU: GPU test suite, bench (C4 headline, reference arm, C1, C3), ncu captures
set -x
timeout 900 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r02_final_gputests.log 2>&1
timeout 500 python bench.py > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_final_bench_c1.json 2> /dev/null
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_c3.json 2> /dev/null
K='regex:^primary_kernel|^shadow_light'
SECS="--section LaunchStats --section Occupancy --section SpeedOfLight --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis"
timeout 600 ncu $SECS --metrics sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k "$K" -s 10 -c 5 -f -o gpurun_out/r02_final_c4 python tools/probe.py nopeak 1000000,3840,2160,4 > gpurun_out/r02_final_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 10 -c 5 -f -o gpurun_out/r02_final_200k python tools/probe.py nopeak 200000,1920,1080,4 > gpurun_out/r02_final_ncu_200k.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cull --no-cpu-baseline --no-e2e > gpurun_out/r02_final_ncu_launch.log 2>&1

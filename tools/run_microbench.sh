#!/bin/bash
# Runs on the GPU box (via gpurun): integer-pipe microbenchmarks + ncu pipe metrics.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.csv
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_micro.csv &
SMI=$!
./tools/pipe_bench 2 512 > gpurun_out/pipe_bench.json
./tools/pipe_bench 1 256 > gpurun_out/pipe_bench_low.json
./tools/fe_bench > gpurun_out/fe_bench.json
kill $SMI
ncu --query-metrics > gpurun_out/ncu_metrics.txt 2>&1
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,smsp__inst_executed.sum \
  --clock-control none --csv --log-file gpurun_out/ncu_fe_bench.csv ./tools/fe_bench > gpurun_out/fe_bench_under_ncu.json 2>&1
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
  --clock-control none --csv --log-file gpurun_out/ncu_pipe_bench.csv ./tools/pipe_bench 2 512 > gpurun_out/pipe_bench_under_ncu.json 2>&1
echo finished

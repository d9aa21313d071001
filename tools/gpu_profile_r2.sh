#!/bin/bash
# GPU box, round 2: plain run first, then the ncu launch list and one --set full capture of each heavy kernel inside the
# default bench (B200_PROFILING.md recipe), plus the SASS op counts the fp64 line of bench.py needs and the lidar's FP32 / L2
# figures the north star asks for.  usage: tools/gpu_profile_r2.sh <tag>
set -x
TAG=${1:-r11}
SETTLE=${SETTLE:-300}
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle $SETTLE --no-cpu-baseline"
SKIP=$(( SETTLE + 5 ))
EXTRA="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s $(( SETTLE * 7 )) -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
# (step_quad_kernel is launched twice per tick: the first launch of the staged solve and its continuation; -c 2 captures both)
for K in step_quad_kernel lidar_kernel drivers_kernel; do
  C=1; [ "$K" = step_quad_kernel ] && C=2
  ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:$K -s $(( SKIP * C )) -c $C -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -n 2 gpurun_out/ncu_${K}_$TAG.log
done
# config 2 (lidar only, 4 096 cars): ns/ray and the same lidar metrics
CMD2="python bench.py --workload lidar --cars 4096 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_lidar_$TAG.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none -k regex:lidar_kernel -s 5 -c 1 -o gpurun_out/prof_lidar4096_$TAG $CMD2 > gpurun_out/ncu_lidar4096_$TAG.log 2>&1
tail -n 1 gpurun_out/plain_$TAG.log | cut -c1-300

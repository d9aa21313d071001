#!/bin/bash
# GPU box: plain run, then ncu launch list + full capture of the heavy kernels (B200_PROFILING.md recipe).
set -x
TAG=${1:-r01}
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 200 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 70 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_quad_kernel -s 205 -c 1 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_quad_resume_kernel -s 205 -c 1 -o gpurun_out/prof_resume_$TAG $CMD > gpurun_out/ncu_resume_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lidar_kernel -s 205 -c 1 -o gpurun_out/prof_lidar_$TAG $CMD > gpurun_out/ncu_lidar_$TAG.log 2>&1
for f in ncu_launch ncu_step ncu_lidar; do tail -n 2 gpurun_out/${f}_$TAG.log; done

#!/bin/bash
TAG=${1:-r01b}
CMD="python bench.py --workload step --cars 65536 --steps 3 --warmup 3 --settle 10 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 12 -c 1 -o gpurun_out/prof_stepw_$TAG $CMD > gpurun_out/ncu_stepw_$TAG.log 2>&1
tail -n 3 gpurun_out/plain_$TAG.log | cut -c1-200

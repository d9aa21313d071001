#!/bin/bash
# GPU box: one ncu --set full capture of the lidar kernel inside the full tick.
TAG=${1:-r01b}
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 200 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:lidar_kernel -s 205 -c 1 -o gpurun_out/prof_lidar_$TAG $CMD > gpurun_out/ncu_lidar_$TAG.log 2>&1
tail -n 1 gpurun_out/ncu_lidar_$TAG.log | cut -c1-100

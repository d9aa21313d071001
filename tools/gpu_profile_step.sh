#!/bin/bash
# GPU box: plain run, then one ncu --set full capture of the vehicle-step kernel inside the full tick.
TAG=${1:-r01b}
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 20 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_.*kernel -s 25 -c 1 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step_$TAG.log 2>&1
tail -n 1 gpurun_out/plain_$TAG.log | cut -c1-200

#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload race --cars 32768 --steps 5 --warmup 2 --settle 300"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:world_step_kernel -s 305 -c 1 -o gpurun_out/prof_world_r13 $CMD > gpurun_out/ncu_world_r13.log 2>&1; tail -n 2 gpurun_out/ncu_world_r13.log

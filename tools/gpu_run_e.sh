#!/bin/bash
# GPU box, round 2 call E: unified step kernel -- sharded bit-identity + lidar off-nominal tests, A/B timing, ncu source-level capture of the step
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_step.py -q -x -k "sharded or graph or single_step or finish or two_fleets or bubble" ) > gpurun_out/gputest_e.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/gputest_e.log
( timeout 300 python -m pytest tests/test_gpu_lidar.py -q -x -k "tilted or multi_car" ) > gpurun_out/gputest_e2.log 2>&1; echo "lidar tests rc=$?"; tail -n 4 gpurun_out/gputest_e2.log
timeout 300 python tools/step_ab.py 2>&1 | grep variant
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 300 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_quad_kernel -s 306 -c 2 -o gpurun_out/prof_step_r11 $CMD > gpurun_out/ncu_step_r11.log 2>&1; tail -n 2 gpurun_out/ncu_step_r11.log

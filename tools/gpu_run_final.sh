#!/bin/bash
# final validation of the round: whole GPU suite, smoke(), the bench lines that get archived, then the profile pass (tag r14)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/gputest_final.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/gputest_final.log | cut -c1-300
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final_tick.json 2> gpurun_out/bench_final_tick.err; echo "bench rc=$?"
timeout 300 python bench.py --workload lidar --cars 4096 > gpurun_out/bench_final_lidar.json 2> gpurun_out/bench_final_lidar.err; echo "lidar rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "ref rc=$?"
for f in tick lidar reference; do tail -n 1 gpurun_out/bench_final_$f.json | cut -c1-300; done
timeout 1500 bash tools/gpu_profile_r2.sh r14 2>&1 | tail -n 6
CMD3="python bench.py --workload race --cars 32768 --steps 3 --warmup 3 --settle 300 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:world_step_kernel -s 305 -c 1 -o gpurun_out/prof_world_r14 $CMD3 > gpurun_out/ncu_world_r14.log 2>&1; tail -n 1 gpurun_out/ncu_world_r14.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lidar_kernel -s 305 -c 1 -o gpurun_out/prof_lidarmulti_r14 $CMD3 > gpurun_out/ncu_lidarmulti_r14.log 2>&1; tail -n 1 gpurun_out/ncu_lidarmulti_r14.log

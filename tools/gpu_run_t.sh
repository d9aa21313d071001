#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_step.py -q -x -k "fused_tick_of_coupled or car_car or sharded or graph or finish" ) > gpurun_out/gputest_t.log 2>&1; echo "step tests rc=$?"; tail -n 3 gpurun_out/gputest_t.log | cut -c1-400
( timeout 300 python -m pytest tests/test_gpu_lidar.py -q -x ) > gpurun_out/gputest_t2.log 2>&1; echo "lidar tests rc=$?"; tail -n 1 gpurun_out/gputest_t2.log
for CARS in 32768 262144; do
timeout 400 python bench.py --workload race --cars $CARS --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_t_race_$CARS.json 2> gpurun_out/bench_t_race.err; echo "race rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_t_race_$CARS.json').read().strip().splitlines()[-1]); print($CARS, d['ms_per_step'], d['value'])"
done
timeout 300 python tools/world_iters.py 2>&1 | cut -c1-120 | tail -8

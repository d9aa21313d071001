#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_fullsize.py -q -x ) > gpurun_out/gputest_q1.log 2>&1; echo "lidar tests rc=$?"; tail -n 2 gpurun_out/gputest_q1.log | cut -c1-300
( timeout 600 python -m pytest tests/test_gpu_step.py -q -x -k "car_car or sharded or graph" ) > gpurun_out/gputest_q2.log 2>&1; echo "step tests rc=$?"; tail -n 2 gpurun_out/gputest_q2.log | cut -c1-300
for CARS in 32768 262144; do
timeout 400 python bench.py --workload race --cars $CARS --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_q_race_$CARS.json 2> gpurun_out/bench_q_race.err; echo "race rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_q_race_$CARS.json').read().strip().splitlines()[-1]); print($CARS, d['ms_per_step'], d['value'], d['episode']['cars_in_coupled_worlds_last_tick_this_rank'])"
done
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q_tick.json 2> gpurun_out/bench_q_tick.err; echo "tick rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_q_tick.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['kernels'], d.get('fp64'))"

#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_step.py -q -x -s -k "car_car or sharded" ) > gpurun_out/gputest_l.log 2>&1; echo "tests rc=$?"; grep -n "report\|passed\|failed\|Error" gpurun_out/gputest_l.log | tail -n 4 | cut -c1-300
for CARS in 32768 262144; do
timeout 400 python bench.py --workload race --cars $CARS --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_l_race_$CARS.json 2> gpurun_out/bench_l_race.err; echo "race rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_l_race_$CARS.json').read().strip().splitlines()[-1]); print($CARS, d['ms_per_step'], d['value'], d['episode']['cars_in_coupled_worlds_last_tick_this_rank'])"
done

#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_step.py -q -x -s ) > gpurun_out/gputest_j.log 2>&1; echo "tests rc=$?"; grep -n "report\|passed\|failed\|Error" gpurun_out/gputest_j.log | tail -n 8
for CARS in 32768 262144; do
timeout 400 python bench.py --workload race --cars $CARS --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_j_race_$CARS.json 2> gpurun_out/bench_j_race.err; echo "race rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_j_race_$CARS.json').read().strip().splitlines()[-1]); print($CARS, d['ms_per_step'], d['value'], d['episode']['cars_in_coupled_worlds_last_tick_this_rank'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'], d['e2e']['value'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_ab.py 2>&1 | grep variant
( timeout 600 python -m pytest tests/test_gpu_step.py -q -x -k "car_car or bubble or flatten" ) > gpurun_out/gputest_h.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/gputest_h.log
timeout 400 python bench.py --workload race --cars 32768 --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_h_race.json 2> gpurun_out/bench_h_race.err; echo "race rc=$?"; cut -c1-200 gpurun_out/bench_h_race.json; tail -n 3 gpurun_out/bench_h_race.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_h_race.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['episode'])"

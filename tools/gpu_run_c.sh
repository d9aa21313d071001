#!/bin/bash
# GPU box, round 2 call C: shard debug, GPU tests of the step / lidar / full-size files (all of them, no -x), bench, lidar mismatch dump
mkdir -p gpurun_out
python tools/debug_shard.py > gpurun_out/debug_shard.log 2>&1; tail -n 12 gpurun_out/debug_shard.log
( time python -m pytest tests/test_gpu_step.py tests/test_gpu_lidar.py tests/test_gpu_fullsize.py -q ) > gpurun_out/gputest_c.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/gputest_c.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'])"
python tools/lidar_mismatch.py > gpurun_out/lidar_mismatch.log 2>&1; tail -n 8 gpurun_out/lidar_mismatch.log

#!/bin/bash
# GPU box, round 2 call D: every pytest invocation under its own timeout (a hung kernel must not eat the call's budget)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_step.py -q -x ) > gpurun_out/gputest_d_step.log 2>&1; echo "step tests rc=$?"
tail -n 12 gpurun_out/gputest_d_step.log
( time timeout 400 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_fullsize.py tests/test_gpu_bench_contract.py -q ) > gpurun_out/gputest_d_rest.log 2>&1; echo "other tests rc=$?"
tail -n 6 gpurun_out/gputest_d_rest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'])"
timeout 300 python tools/lidar_mismatch.py > gpurun_out/lidar_mismatch.log 2>&1; tail -n 8 gpurun_out/lidar_mismatch.log

#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/step_ab.py 2>&1 | grep variant
( timeout 900 python -m pytest tests/test_gpu_step.py -q -x ) > gpurun_out/gputest_g.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/gputest_g.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_g.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'], d['e2e']['value'])"

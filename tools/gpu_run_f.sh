#!/bin/bash
mkdir -p gpurun_out
rm -f variants/*.so
timeout 200 python tools/step_ab.py 2>&1 | grep variant
( timeout 900 python -m pytest tests/test_gpu_step.py -q -x ) > gpurun_out/gputest_f.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/gputest_f.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_f.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'], d['e2e']['value'])"
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 300 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_quad_kernel -s 306 -c 2 -o gpurun_out/prof_step_r12 $CMD > gpurun_out/ncu_step_r12.log 2>&1; tail -n 1 gpurun_out/ncu_step_r12.log

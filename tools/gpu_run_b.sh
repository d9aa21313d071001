#!/bin/bash
# GPU box, round 2 call B: step/contact tests after the new contact rules, graph A/B, bench line.
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -x -q ) > gpurun_out/gputest_b.log 2>&1; echo "pytest rc=$?"
tail -n 8 gpurun_out/gputest_b.log
python tools/graph_ab.py > gpurun_out/graph_ab.json 2> gpurun_out/graph_ab.err; echo "graph rc=$?"; cat gpurun_out/graph_ab.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_b.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'])"

#!/bin/bash
# GPU box, round 2 call A: GPU test suite, default bench line, regime sweep, compute-sanitizer logs of the smoke.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_a.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/gputest_a.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --settle 200 --no-cpu-baseline > gpurun_out/bench_a_settle200.json 2>> gpurun_out/bench_a.err
python tools/regime_sweep.py > gpurun_out/regime_sweep.json 2> gpurun_out/regime.err; echo "sweep rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  tail -n 4 gpurun_out/sanitizer_$tool.log
done
cut -c1-400 gpurun_out/bench_a.json; cat gpurun_out/regime_sweep.json

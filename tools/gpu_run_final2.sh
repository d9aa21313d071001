#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/gputest_final2.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/gputest_final2.log | cut -c1-300
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_final2.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final2_tick.json 2> gpurun_out/bench_final2_tick.err; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_final2_tick.json | cut -c1-260
CMD="python bench.py --workload tick --cars 65536 --steps 3 --warmup 3 --settle 300 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 80 --csv --log-file gpurun_out/launches_r15.csv $CMD > gpurun_out/ncu_launch_r15.log 2>&1; echo "launch list rc=$?"

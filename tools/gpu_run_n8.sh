#!/bin/bash
# 8 GPUs of one box: weak-scaling tick, race worlds (config 5), sharded episode to completion (config 4)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 400 $TR bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8_tick.json 2> gpurun_out/bench_n8_tick.err; echo "tick rc=$?"
tail -n 1 gpurun_out/bench_n8_tick.json | cut -c1-200
timeout 400 $TR bench.py --gpus 8 --workload race --cars 262144 --steps 200 --warmup 5 --settle 300 > gpurun_out/bench_n8_race.json 2> gpurun_out/bench_n8_race.err; echo "race rc=$?"
tail -n 1 gpurun_out/bench_n8_race.json | cut -c1-200
timeout 600 $TR bench.py --gpus 8 --workload episode --full-episode --lap-target 1 --cars 1048576 --warmup 5 > gpurun_out/bench_n8_episode_full.json 2> gpurun_out/bench_n8_episode.err; echo "episode rc=$?"
tail -n 1 gpurun_out/bench_n8_episode_full.json | cut -c1-200

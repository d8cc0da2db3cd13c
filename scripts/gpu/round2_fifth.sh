#!/bin/bash
# GPU call 5 of round 2: branch-free flat replay refill, guard-band tests, full bench line with the new legs
T=r04e
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/replay_bench.py 4194304 12 4096 | tee gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 4194304 12 1 | tee -a gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 16777216 12 4096 | tee -a gpurun_out/${T}_replay_bench.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench_1gpu.err; cut -c1-150 gpurun_out/${T}_bench_1gpu.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_replay_fast python scripts/replay_bench.py 4194304 4 4096 > gpurun_out/${T}_ncu_replay.log 2>&1; tail -1 gpurun_out/${T}_ncu_replay.log | cut -c1-100

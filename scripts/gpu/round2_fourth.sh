#!/bin/bash
# GPU call 4 of round 2: flat refill loop of the replay common-case kernel; out-of-line single crossings (variant)
T=r04d
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/quick_bench.sh libecg.so
python scripts/replay_bench.py 4194304 12 4096 | tee gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 4194304 12 1 | tee -a gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 16777216 12 4096 | tee -a gpurun_out/${T}_replay_bench.jsonl
L=$PWD/element-crush-gym_b200/lib
ECG_LIB=$L/libecg_s9_cross.so python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e --no-extra-legs 2>&1 | tail -1 > gpurun_out/${T}_s9_cross.json; python -c "
import json; d=json.load(open('gpurun_out/${T}_s9_cross.json')); print('s9_cross', '%.3e' % d['value'], d['roofline']['handed_off_to_exact_kernel'], d['roofline']['avg_launch_ms'], d['roofline']['exact_kernel_avg_ms'])"
python scripts/sweep_shapes.py 4194304 24 6:4,14:7,15:8 2>&1 | tee gpurun_out/${T}_sweep.jsonl | cut -c1-120
ECG_LIB=$L/libecg_s6_cross.so python scripts/sweep_shapes.py 4194304 24 6:4 2>&1 | tail -1 | tee -a gpurun_out/${T}_sweep.jsonl | cut -c1-120
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_replay_fast python scripts/replay_bench.py 4194304 4 4096 > gpurun_out/${T}_ncu_replay.log 2>&1; tail -1 gpurun_out/${T}_ncu_replay.log | cut -c1-100

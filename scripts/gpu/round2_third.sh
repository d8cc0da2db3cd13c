#!/bin/bash
# GPU call 3 of round 2: two-kernel replay step, block-size defaults for the large boards
T=r04c
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/quick_bench.sh libecg.so
python scripts/replay_bench.py 4194304 12 4096 | tee gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 4194304 12 1 | tee -a gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 16777216 12 4096 | tee -a gpurun_out/${T}_replay_bench.jsonl
L=$PWD/element-crush-gym_b200/lib
python scripts/sweep_shapes.py 4194304 24 12:7,13:7,14:7,15:8,16:8 > gpurun_out/${T}_large_sweep.jsonl 2>&1; cut -c1-110 gpurun_out/${T}_large_sweep.jsonl
for v in s16_b256ls:16:8 s16_b320:16:8 s15_b256:15:8 s14_b256:14:7 s14_b384:14:7 s13_b256:13:7 s12_ls:12:7; do
  IFS=: read lib r t <<< "$v"
  ECG_LIB=$L/libecg_$lib.so python scripts/sweep_shapes.py 4194304 24 $r:$t 2>&1 | tail -1 | tee -a gpurun_out/${T}_block_variants.jsonl | cut -c1-110
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_replay_fast python scripts/replay_bench.py 4194304 4 4096 > gpurun_out/${T}_ncu_replay.log 2>&1; tail -1 gpurun_out/${T}_ncu_replay.log | cut -c1-100

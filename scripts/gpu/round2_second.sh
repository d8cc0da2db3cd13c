#!/bin/bash
# GPU call 2 of round 2: micro-optimised 9x9 kernels, block-size variants for 6 / 12 / 16, ncu captures (6x6x4 and
# 16x16x8 common-case kernels, the replay step kernel)
T=r04b
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/quick_bench.sh libecg.so
L=$PWD/element-crush-gym_b200/lib
python scripts/sweep_shapes.py 4194304 24 > gpurun_out/${T}_shape_sweep.jsonl 2>&1; cut -c1-110 gpurun_out/${T}_shape_sweep.jsonl
for v in s16_b256:16:8 s16_b384:16:8 s16_f256:16:8 s12_b256:12:7 s12_b384:12:7 s12_f384:12:7 s6_f640:6:4 s6_f768:6:4; do
  IFS=: read lib r t <<< "$v"
  ECG_LIB=$L/libecg_$lib.so python scripts/sweep_shapes.py 4194304 24 $r:$t 2>&1 | tail -1 | tee -a gpurun_out/${T}_block_variants.jsonl | cut -c1-110
done
python scripts/replay_bench.py 4194304 12 4096 | tee gpurun_out/${T}_replay_bench.jsonl
python scripts/replay_bench.py 4194304 12 1 | tee -a gpurun_out/${T}_replay_bench.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 8 -f -o gpurun_out/${T}_fast_6x6x4 python scripts/sweep_shapes.py 4194304 4 6:4 > gpurun_out/${T}_ncu_6.log 2>&1; tail -1 gpurun_out/${T}_ncu_6.log | cut -c1-100
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 8 -f -o gpurun_out/${T}_fast_16x16x8 python scripts/sweep_shapes.py 4194304 4 16:8 > gpurun_out/${T}_ncu_16.log 2>&1; tail -1 gpurun_out/${T}_ncu_16.log | cut -c1-100
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 4 -f -o gpurun_out/${T}_replay_9x9x6 python scripts/replay_bench.py 4194304 4 4096 > gpurun_out/${T}_ncu_replay.log 2>&1; tail -1 gpurun_out/${T}_ncu_replay.log | cut -c1-100
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-legs > gpurun_out/${T}_ncu_fast.log 2>&1; tail -1 gpurun_out/${T}_ncu_fast.log | cut -c1-100
ls -la gpurun_out/${T}_*

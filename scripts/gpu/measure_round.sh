# usage: bash scripts/gpu/measure_round.sh <tag>   (one GPU) -> gpurun_out/<tag>_*: tests, bench lines, ncu launch list and full captures
T=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${T}_pytest_gpu.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; cut -c1-200 gpurun_out/${T}_bench_1gpu.json
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench_reference_arm.json
# launch list of the same command (short run): per-launch times are cold-cache and serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra-legs > gpurun_out/${T}_ncu_launches.log 2>&1
# full captures of the two step kernels at the bench size
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-legs > gpurun_out/${T}_ncu_fast.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 7 -f -o gpurun_out/${T}_exact python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-legs > gpurun_out/${T}_ncu_exact.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/${T}_replay_fast python scripts/replay_bench.py 16777216 4 4096 > gpurun_out/${T}_ncu_replay.log 2>&1
python scripts/sweep_shapes.py 4194304 24 4:4,5:4,6:4,7:5,8:5,9:6,10:6,11:7,12:7,13:7,14:7,15:8,16:8 > gpurun_out/${T}_shape_sweep.jsonl 2>&1; cut -c1-120 gpurun_out/${T}_shape_sweep.jsonl
python scripts/mcts_bench.py > gpurun_out/${T}_mcts_bench.jsonl 2>&1; tail -1 gpurun_out/${T}_mcts_bench.jsonl | cut -c1-200
python scripts/mcts_bench.py --refill replay --sims 64 >> gpurun_out/${T}_mcts_bench.jsonl 2>&1; tail -1 gpurun_out/${T}_mcts_bench.jsonl | cut -c1-200
python scripts/rollout_bench.py > gpurun_out/${T}_rollout_bench.txt 2>&1; tail -2 gpurun_out/${T}_rollout_bench.txt | cut -c1-200
python scripts/replay_bench.py 16777216 12 4096 > gpurun_out/${T}_replay_bench.jsonl; python scripts/replay_bench.py 16777216 12 1 >> gpurun_out/${T}_replay_bench.jsonl; cut -c1-160 gpurun_out/${T}_replay_bench.jsonl

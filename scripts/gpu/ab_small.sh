#!/bin/bash
# usage: scripts/gpu/ab_small.sh lib1.so lib2.so ...  -> step rate at 2^24 .. 2^17 boards and rollouts at 2^17 / 2^20, per library
for L in "$@"; do
  export ECG_LIB=$PWD/element-crush-gym_b200/lib/$L
  for B in 16777216 4194304 2097152 524288 131072; do
    python bench.py --boards $B --steps 24 --warmup 4 --no-cpu-baseline --no-e2e --no-extra-legs 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('$L', $B, '%.3e' % d['value'], 'step %.3f ms' % d['ms_per_step'], 'fast %.3f' % r['avg_launch_ms'], 'exact %.3f' % r['exact_kernel_avg_ms'])"
  done
  for B in 131072 1048576; do python scripts/rollout_bench.py $B | tail -1; done
done

#!/bin/bash
# usage: scripts/gpu/ab_exact.sh lib1.so lib2.so ...  -> step rate and exact-kernel time per library variant, at 2^24 and 2^22 boards
for L in "$@"; do for B in 16777216 4194304; do
  ECG_LIB=$PWD/element-crush-gym_b200/lib/$L python bench.py --boards $B --steps 24 --warmup 4 --no-cpu-baseline --no-e2e --no-extra-legs 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('$L', $B, '%.3e' % d['value'], 'step %.3f ms' % d['ms_per_step'], 'fast %.3f' % r['avg_launch_ms'], 'exact %.3f' % r['exact_kernel_avg_ms'])"
done; done

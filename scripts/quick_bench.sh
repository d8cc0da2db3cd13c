#!/bin/bash
# usage: scripts/quick_bench.sh lib1.so lib2.so ...  -> kernel-only env-steps/s per library variant
for L in "$@"; do
  ECG_LIB=$PWD/element-crush-gym_b200/lib/$L python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$L', '%.3e' % d['value'], '%.3f ms' % d['ms_per_step'])"
done

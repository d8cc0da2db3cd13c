# usage: scripts/gpu/ab_pool.sh lib1.so lib2.so ...  (pooled-kernel variants holding only size 9: parity tests of the pooled kernel, then the kernel-only bench)
for L in "$@"; do
  echo "== $L"
  ECG_LIB=$PWD/element-crush-gym_b200/lib/$L timeout 300 python -m pytest tests/test_gpu_pool.py -m gpu -x -q 2>&1 | tail -1
  ECG_LIB=$PWD/element-crush-gym_b200/lib/$L timeout 120 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e --no-extra-legs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$L', '%.3e' % d['value'], '%.3f ms' % d['ms_per_step'], 'common-case %.3f ms' % d['roofline']['avg_launch_ms'])"
done

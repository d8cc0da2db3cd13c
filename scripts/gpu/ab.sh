# usage: scripts/gpu/ab.sh lib1.so lib2.so ...   (A/B kernel-only bench of library variants; gpu tests on the default lib first)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/quick_bench.sh "$@"

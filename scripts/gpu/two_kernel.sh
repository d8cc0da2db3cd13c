python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/quick_bench.sh libecg.so
ECG_SINGLE_KERNEL=1 bash scripts/quick_bench.sh libecg.so
python bench.py --steps 24 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e e2e %.3e launches %s' % (d['value'], d['e2e']['value'], d['gpu_launches']))"

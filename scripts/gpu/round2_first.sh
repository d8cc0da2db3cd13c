#!/bin/bash
# first GPU call of round 2: parity tests, bench (1 GPU), reference arm, compute-sanitizer on a small batch
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r04_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r04_pytest_gpu.log
tail -5 gpurun_out/r04_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r04_bench_1gpu.json 2> gpurun_out/r04_bench_1gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r04_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r04_bench_reference_arm.json 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r04_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r04_sanitizer_$tool.log
done

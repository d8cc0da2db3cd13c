set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_quick.log 2>&1
ECG_LIB=$PWD/element-crush-gym_b200/lib/libecg_phases.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 3 -o gpurun_out/r02_phases python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_phases.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 3 -o gpurun_out/r02_full python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_pytest.log; tail -1 gpurun_out/r02_bench_quick.log

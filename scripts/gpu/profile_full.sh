# usage: scripts/gpu/profile_full.sh <lib.so> <tag>  -> gpurun_out/<tag>.ncu-rep (--set full, source) of the step kernel
ECG_LIB=$PWD/element-crush-gym_b200/lib/$1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 3 -f -o gpurun_out/$2 python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log

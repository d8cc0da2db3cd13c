# usage: scripts/gpu/profile_fast.sh <tag>  -> gpurun_out/<tag>_fast.ncu-rep / <tag>_exact.ncu-rep (--set full, source) of the two step kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 6 -f -o gpurun_out/$1_fast python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/$1_fast.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lane_kernel -c 1 -s 7 -f -o gpurun_out/$1_exact python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/$1_exact.log 2>&1
tail -1 gpurun_out/$1_exact.log | cut -c1-100

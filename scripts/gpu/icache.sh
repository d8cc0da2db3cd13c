# usage: scripts/gpu/icache.sh lib1.so ...  -> kernel-only bench + instruction-cache counters of the step kernel per library variant
for L in "$@"; do
  bash scripts/quick_bench.sh $L
  ECG_LIB=$PWD/element-crush-gym_b200/lib/$L ncu --clock-control none -k regex:lane_kernel -c 1 -s 3 \
    --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
    python bench.py --boards 4194304 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | grep -E "icc_|gcc__|inst_executed|inst_issued|time_duration" | sed "s/^/   /"
done

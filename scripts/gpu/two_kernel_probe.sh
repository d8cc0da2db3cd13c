python - <<'PY'
import importlib, torch, sys
sys.path.insert(0, '.')
E = importlib.import_module("element-crush-gym_b200")
n = 1 << 22
env = E.BatchedMatch3Env(n, 9, 9, 6, num_moves=1 << 30, seed=12345, refill="philox")
b = env.board
for i in range(6):
    b.apply_action(None)
    torch.cuda.synchronize()
    print("step", i, "handed off", int(b._scratch[0].item()), "of", n, "= %.2f%%" % (100.0 * b._scratch[0].item() / n))
PY
ncu --clock-control none -k regex:lane_kernel -c 6 -s 6 --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__inst_issued.avg.pct_of_peak_sustained_active python bench.py --boards 4194304 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "lane_kernel|time_duration|inst_executed|icc_request|inst_issued" | cut -c1-150

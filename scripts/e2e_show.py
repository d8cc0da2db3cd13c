import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("e2e","e2e_uint8_direct","e2e_uint8_host_expand","e2e_nibbles"):
    e=d[k]; print(k, "%.3e"%e["value"], "%.2f ms"%e["ms_per_step"], e.get("host_expand_threads"), "%.1f GB/s"%e["pcie_d2h_gb_s_per_gpu"])
print(d["value"])

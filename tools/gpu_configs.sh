#!/bin/bash
# BASELINE.json configs C3/C4/C5 on one GPU with the current library (name of the result set = $1, default r1d)
n=${1:-r1d}
python tools/run_configs.py --configs C3,C4,C5 > gpurun_out/${n}_configs.jsonl 2> gpurun_out/${n}_configs.err
VSRT_BENCH_MODE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${n}_bench_dfs.json 2>> gpurun_out/${n}_configs.err
python - <<PY
import json
f = "gpurun_out/${n}_configs.jsonl"
for l in open(f):
    d = json.loads(l)
    print("  ", d.get("config"), d.get("bounce", d.get("budget", d.get("summary", ""))), "%.1f Mrays/s" % (d.get("rays_per_s", 0) / 1e6), "k1 %.2f" % d.get("k1_ms", 0), "k3 %.2f" % d.get("k3_ms", 0))
d = json.loads(open("gpurun_out/${n}_bench_dfs.json").read().strip().splitlines()[-1])
print("DFS bench: %.1f Mrays/s" % (d["value"] / 1e6), d["roofline"]["step_breakdown_ms"], "bytes/ray %.0f" % d["config"]["bytes_per_ray"])
PY

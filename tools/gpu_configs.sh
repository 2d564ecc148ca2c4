#!/bin/bash
# BASELINE.json configs C3/C4/C5 on one GPU with the current library; C4 again with the L1 next-node prefetch on
python tools/run_configs.py --configs C3,C4,C5 > gpurun_out/r1b_configs.jsonl 2> gpurun_out/r1b_configs.err
VSRT_PREFETCH=1 python tools/run_configs.py --configs C4 > gpurun_out/r1b_configs_c4_prefetch.jsonl 2>> gpurun_out/r1b_configs.err
python - <<'PY'
import json
for f in ("gpurun_out/r1b_configs.jsonl", "gpurun_out/r1b_configs_c4_prefetch.jsonl"):
    print(f)
    for l in open(f):
        d = json.loads(l)
        print("  ", d.get("config"), d.get("bounce", d.get("budget", d.get("summary", ""))), "%.1f Mrays/s" % (d.get("rays_per_s", 0) / 1e6), "k1 %.2f" % d.get("k1_ms", 0), "k3 %.2f" % d.get("k3_ms", 0))
PY

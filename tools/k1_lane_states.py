#!/usr/bin/env python
"""Where the lanes of K1's warps are, per inner round, on the bench workload.  Needs a library built with
-DVSRT_K1_STATS=1:  VSRT_NVCC_EXTRA=-DVSRT_K1_STATS=1 python -c "import __graft_entry__ as g; g.build_cuda(force=True)"."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import vsrt.api as api
from vsrt import scene as sc, _abi

mode = int(os.environ.get("VSRT_BENCH_MODE", "1"))
s = sc.Scene(1_000_000, seed=0x5EED0001 + 1)
rays = sc.rays_primary(1920, 1080, flags=(_abi.FLAG_OPAQUE if mode == 0 else 0))
ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
L = ctx.L
out = (ctypes.c_ulonglong * 16)()
ctx.trace_device(mode, rd.data_ptr(), len(rays))
L.vsrt_debug_k1_stats(out)                      # clear after the warm-up batch
ctx.trace_device(mode, rd.data_ptr(), len(rays))
assert L.vsrt_debug_k1_stats(out) == 0
v = [int(x) for x in out]
names = ["idle", "defer", "finished", "pop", "internal", "instance", "leaf"]
rounds = v[0]
print(json.dumps({"inner_rounds": rounds, "rounds_with_internal_phase": v[9],
                  "lanes_per_round": {n: round(v[1 + i] / rounds, 2) for i, n in enumerate(names)},
                  "leaf_phases": v[10], "lanes_per_leaf_phase": round(v[11] / max(v[10], 1), 2),
                  "refills": v[12], "idle_lanes_per_refill": round(v[13] / max(v[12], 1), 2)}))

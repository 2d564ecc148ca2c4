#!/usr/bin/env python
"""Per-ray peak traversal-stack depth (node entries; both lists / `current` alone) on the bench workload or C3 bounce rays.
Needs the stats build: VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_stats.so (tools/build_variant.sh stats -DVSRT_K1_STATS=1)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import vsrt.api as api
from vsrt import scene as sc, _abi

which = sys.argv[1] if len(sys.argv) > 1 else "bench"
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 512
if which == "bench":
    s = sc.Scene(1_000_000, seed=0x5EED0001 + 1); rays = sc.rays_primary(1920, 1080, flags=0)
else:
    s = sc.Scene(2_000_000, seed=0x5EED0001 + 2, kind=sc.CLUSTERED); rays = sc.rays_primary(1920, 1080, spp=2, flags=0)
ctx = api.Context(max_treelet_size=budget, device=0); ctx.register(s); ctx.form_treelets()
if which != "bench":
    g = ctx.trace(1, rays, want_trace=False); rays = s.bounce(rays, g["hits"], 77, 0, 0)
rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
out = (ctypes.c_ulonglong * 128)()
ctx.L.vsrt_debug_k1_depth(out)
ctx.trace_device(1, rd.data_ptr(), len(rays))
assert ctx.L.vsrt_debug_k1_depth(out) == 0
v = np.array([int(x) for x in out], dtype=np.int64)
both, cur = v[:64], v[64:]
n = both.sum()
cum = np.cumsum(both) / max(n, 1)
print(json.dumps({"workload": which, "budget": budget, "rays": int(n), "peak_both_lists_hist": both[:40].tolist(), "peak_current_hist": cur[:24].tolist(),
                  "frac_rays_peak_le": {str(k): round(float(cum[k]), 5) for k in (4, 6, 8, 10, 12, 16, 20, 24)}}))

#!/usr/bin/env python
"""Latency of the 32-lane call vsrt_trace_ray_warp (the drop-in for one trace_ray warp instruction) and of small batches."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np
import __graft_entry__ as g
import vsrt.api as api
from vsrt import scene as sc, _abi

s = sc.Scene(1_000_000, seed=0x5EED0001 + 1)
rays = sc.rays_primary(1920, 1080)
ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
mid = len(rays) // 2
for n in (32, 1024, 32768, 1 << 20):
    batch = rays[mid:mid + n]
    for _ in range(3):
        ctx.trace(_abi.MODE_TREELET, batch) if n != 32 else ctx.trace_warp(batch)
    reps = 200 if n <= 1024 else 20 if n <= 32768 else 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.trace(_abi.MODE_TREELET, batch) if n != 32 else ctx.trace_warp(batch)
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"call": "vsrt_trace_ray_warp" if n == 32 else "vsrt_trace_rays", "rays": n, "us_per_call": dt * 1e6, "rays_per_s": n / dt}), flush=True)

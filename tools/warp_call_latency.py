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

# the C-ABI call alone: buffers allocated once, ctypes call in the loop (what a C++ caller pays)
import ctypes
L = ctx.L
r32 = np.ascontiguousarray(rays[mid:mid + 32], _abi.RAY)
hits = np.zeros(32, _abi.HIT); counts = np.zeros(32, np.uint32); txns = np.zeros(32 * 1024, _abi.TXN); total = ctypes.c_uint64()
args = (ctx.h, ctx.tlas, 0xffffffff, _abi.ptr(r32), _abi.ptr(hits), _abi.ptr(counts), _abi.ptr(txns), len(txns), ctypes.byref(total))
for _ in range(20):
    assert L.vsrt_trace_ray_warp(*args) == 0
t0 = time.perf_counter()
for _ in range(500):
    L.vsrt_trace_ray_warp(*args)
dt = (time.perf_counter() - t0) / 500
print(json.dumps({"call": "vsrt_trace_ray_warp (preallocated buffers)", "rays": 32, "us_per_call": dt * 1e6, "rays_per_s": 32 / dt, "records": int(total.value)}), flush=True)
for n in (1024, 4096):
    rn = np.ascontiguousarray(rays[mid:mid + n], _abi.RAY)
    h = np.zeros(n, _abi.HIT); offs = np.zeros(n + 1, np.uint64); tx = np.zeros(n * 64, _abi.TXN); tid = np.zeros(n * 64, np.uint64)
    a = (ctx.h, ctx.tlas, _abi.MODE_TREELET, n, _abi.ptr(rn), _abi.ptr(h), _abi.ptr(offs), _abi.ptr(tx), len(tx), _abi.ptr(tid), ctypes.byref(total))
    for _ in range(10):
        assert L.vsrt_trace_rays(*a) == 0
    t0 = time.perf_counter()
    for _ in range(200):
        L.vsrt_trace_rays(*a)
    dt = (time.perf_counter() - t0) / 200
    print(json.dumps({"call": "vsrt_trace_rays (preallocated buffers)", "rays": n, "us_per_call": dt * 1e6, "rays_per_s": n / dt, "records": int(total.value)}), flush=True)


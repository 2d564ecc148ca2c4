#!/usr/bin/env python
"""Times the RT-unit replay helpers (vsrt_sort_trace, vsrt_prefetch_vote, vsrt_prefetch_chunks) on the bench workload
(1M triangles, 1080p primary rays, 512 B treelets) and the CPU oracle on a sample of it; one JSON line each."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import __graft_entry__ as g
g.build_cpu()
import vsrt.api as api
from vsrt import scene as sc, _abi
import oracles

n_tri = int(os.environ.get("VSRT_REPLAY_TRI", "1000000"))
s = sc.Scene(n_tri, seed=0x5EED0001 + 1)
rays = sc.rays_primary(1920, 1080)
ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
n_txn = ctx.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))

def timed(f, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best, r

for method in (1, 0):
    dt, _ = timed(lambda: ctx._ck(ctx.L.vsrt_sort_trace(ctx.h, method)))
    print(json.dumps({"helper": "vsrt_sort_trace", "method": method, "rays": len(rays), "records": n_txn, "ms": dt * 1e3, "records_per_s": n_txn / dt}), flush=True)
go = np.arange(0, len(rays) + 1, 32, dtype=np.uint64)
for h in (0, 2):
    dt, dec = timed(lambda: ctx.prefetch_vote(go, h, 0.75))
    print(json.dumps({"helper": "vsrt_prefetch_vote", "heuristic": h, "groups": len(go) - 1, "rays": len(rays), "ms": dt * 1e3, "groups_per_s": (len(go) - 1) / dt}), flush=True)
    dt, (offs, ca, co) = timed(lambda: ctx.prefetch_chunks(dec, h))
    print(json.dumps({"helper": "vsrt_prefetch_chunks", "heuristic": h, "groups": len(go) - 1, "chunks": int(len(ca)), "ms": dt * 1e3}), flush=True)
# CPU oracle on a sample
orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
orc.register(s); orc.form(512)
sub = rays[::64][:16384]
t = orc.trace(1, sub)
for method in (1, 0):
    t0 = time.perf_counter(); orc.sort_trace(method, t); dt = time.perf_counter() - t0
    print(json.dumps({"helper": "cpu sort_mem_accesses (%s)" % orc.kind, "method": method, "rays": len(sub), "records": int(len(t["txns"])), "ms": dt * 1e3, "records_per_s": len(t["txns"]) / dt}), flush=True)
t0 = time.perf_counter()
for g0 in range(0, len(sub), 32):
    orc.prefetch_vote(t, np.arange(g0, min(g0 + 32, len(sub))), 0)
dt = time.perf_counter() - t0
print(json.dumps({"helper": "cpu prefetch vote block (%s)" % orc.kind, "groups": (len(sub) + 31) // 32, "ms": dt * 1e3, "groups_per_s": ((len(sub) + 31) // 32) / dt}), flush=True)

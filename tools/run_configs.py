#!/usr/bin/env python
"""Runs the BASELINE.json configs other than the bench headline on one GPU and prints one JSON line per
measurement (kept under profiles/): they are parity/scale cases, not bench lines.

  C3  2M-triangle clustered scene, 1080p primary + 4 diffuse bounces (incoherent), TREELET mode
  C4  10M-triangle scene, treelet budget sweep 512 B .. 48 KB: formation time + traversal rays/s per budget
  C5  10M-triangle scene, 4K x 8 spp primary rays traced in batches (one GPU's shard when --shard r/N is given)

Every measurement is device-resident (rays in HBM), CUDA-event timed inside the library (K1 + scan + K3)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from vsrt import scene as sc, _abi, shard  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def trace_batch(ctx, mode, rays, reps=3):
    """Device-resident trace of one batch; returns (ms per pass [K1+scan+K3], device results, hits on host)."""
    dev = torch.device("cuda", 0)
    rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    best = None
    for _ in range(reps):
        ctx.trace_device(mode, rd.data_ptr(), len(rays))
        r = ctx.device_results()
        ms = r.traverse_ms + r.scan_ms + r.compact_ms
        if best is None or ms < best[0]:
            best = (ms, r.traverse_ms, r.compact_ms, r.n_txn, r.algorithmic_bytes)
    hits = np.zeros(len(rays), _abi.HIT)
    torch.cuda.synchronize()
    import ctypes
    ctypes.memmove  # hits are fetched through torch to stay off the library's stream
    h = torch.empty(len(rays) * _abi.HIT.itemsize, dtype=torch.uint8, device=dev)
    res = ctx.device_results()
    src = torch.as_tensor(_Dev(res.hits, len(rays) * _abi.HIT.itemsize), device=dev)
    h.copy_(src)
    hits = h.cpu().numpy().view(_abi.HIT)
    return best, hits


class _Dev:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def c3(api):
    s = sc.Scene(2_000_000, seed=0x5EED0001 + 2, kind=sc.CLUSTERED)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s)
    ti = ctx.form_treelets()
    rays = sc.rays_primary(1920, 1080, flags=0)
    tot_rays, tot_ms = 0, 0.0
    for bounce in range(5):
        (ms, k1, k3, ntxn, ab), hits = trace_batch(ctx, _abi.MODE_TREELET, rays)
        emit(config="C3", bounce=bounce, rays=len(rays), ms=ms, k1_ms=k1, k3_ms=k3, rays_per_s=len(rays) / ms * 1e3,
             records_per_ray=ntxn / max(len(rays), 1), bytes_per_ray=ab / max(len(rays), 1), hit_fraction=float(hits["hit_geometry"].mean()),
             treelets=int(ti.n_treelets), form_ms=ti.form_ms, triangles=2_000_000)
        if bounce > 0:
            tot_rays += len(rays); tot_ms += ms
        rays = s.bounce(rays, hits, 77, bounce, 0)
        if len(rays) == 0:
            break
    emit(config="C3", summary="bounces 1-4 (incoherent)", rays=tot_rays, ms=tot_ms, rays_per_s=tot_rays / max(tot_ms, 1e-9) * 1e3)
    ctx.close()


def c4(api, n_tri):
    s = sc.Scene(n_tri, seed=0x5EED0001 + 3)
    prim = sc.rays_primary(1920, 1080, flags=0)
    ctx0 = api.Context(max_treelet_size=512, device=0)
    ctx0.register(s); ctx0.form_treelets()
    (_, _, _, _, _), hits = trace_batch(ctx0, _abi.MODE_TREELET, prim, reps=1)
    rays = s.bounce(prim, hits, 5, 1, 0)          # C3-style incoherent rays
    ctx0.close()
    for budget in (512, 1024, 2048, 4096, 8192, 16384, 32768, 49152):
        ctx = api.Context(max_treelet_size=budget, device=0)
        ctx.register(s)
        t0 = time.time(); ti = ctx.form_treelets(); wall = time.time() - t0
        (ms, k1, k3, ntxn, ab), _ = trace_batch(ctx, _abi.MODE_TREELET, rays)
        emit(config="C4", triangles=n_tri, budget=budget, treelets=int(ti.n_treelets), list_entries=int(ti.n_list_entries), form_ms=ti.form_ms, form_wall_s=wall,
             rays=len(rays), ms=ms, k1_ms=k1, k3_ms=k3, rays_per_s=len(rays) / ms * 1e3, records_per_ray=ntxn / len(rays), bytes_per_ray=ab / len(rays))
        ctx.close()


def c5(api, n_tri, shard_spec, batch):
    r, n = (int(x) for x in shard_spec.split("/"))
    s = sc.Scene(n_tri, seed=0x5EED0001 + 4)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ti = ctx.form_treelets()
    W, H, spp = 3840, 2160, 8
    first, count = shard.shard_range(W * H * spp, n, r)
    done, ms_tot, rec = 0, 0.0, 0
    while done < count:
        m = min(batch, count - done)
        rays = sc.rays_primary(W, H, spp=spp, seed=9, first=first + done, count=m)
        (ms, k1, k3, ntxn, ab), _ = trace_batch(ctx, _abi.MODE_TREELET, rays, reps=2)   # best of 2: the first pass of a batch size grows the staging buffers
        done += m; ms_tot += ms; rec += ntxn
    c = ctx.counters()
    emit(config="C5", shard=shard_spec, triangles=n_tri, rays=count, batches=(count + batch - 1) // batch, ms=ms_tot, rays_per_s=count / ms_tot * 1e3,
         records_per_ray=rec / count, treelets=int(ti.n_treelets), form_ms=ti.form_ms, ray_count=c["ray_count"] // 2, hits=c["num_hits"] // 2,
         max_nodes_per_ray=c["max_nodes_per_ray"], max_tree_depth=c["max_tree_depth"])
    assert c["ray_count"] == 2 * count and sum(c["mem_access_type_%d" % i] for i in range(9)) == 2 * rec   # two passes per batch
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C3,C4,C5")
    ap.add_argument("--c4-triangles", type=int, default=10_000_000)
    ap.add_argument("--c5-triangles", type=int, default=10_000_000)
    ap.add_argument("--shard", default="0/8", help="C5: rank/world of the 4K x 8spp frame traced here")
    ap.add_argument("--batch", type=int, default=4_147_200)
    a = ap.parse_args()
    g.build()
    import vsrt.api as api
    for c in a.configs.split(","):
        if c == "C3":
            c3(api)
        elif c == "C4":
            c4(api, a.c4_triangles)
        elif c == "C5":
            c5(api, a.c5_triangles, a.shard, a.batch)


if __name__ == "__main__":
    main()

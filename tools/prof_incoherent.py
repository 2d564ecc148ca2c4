#!/usr/bin/env python
"""Traces one incoherent batch of BASELINE.json configs[2] / configs[3] a few times so that ncu can capture K1 on it:

    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_traverseILi1ELi96ELb0 \
        -s <skip> -c 1 -o gpurun_out/prof_k1_C3 python tools/prof_incoherent.py --config C3

C3: 2 M-triangle clustered scene, bounce-1 rays of a 1080p x 4 spp frame.  C4: 10 M-triangle scene, bounce-1 rays of a 1080p
frame.  The hot K1 launches are: 1 (primary rays) + --reps (the incoherent batch): skip 1 to land on the first of them.
Prints one JSON line with the library's CUDA-event stage times (not a bench value when run under ncu)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from vsrt import scene as sc, _abi  # noqa: E402


class _Dev:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--budget", type=int, default=512)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--triangles", type=int, default=0)
    a = ap.parse_args()
    g.build()
    os.environ.setdefault("VSRT_STAGE_CAP", "512")   # no batch is redone because a ray outgrew its staging segment: launch counts are fixed
    import vsrt.api as api
    dev = torch.device("cuda", 0)
    if a.config == "C3":
        s = sc.Scene(a.triangles or 2_000_000, seed=0x5EED0001 + 2, kind=sc.CLUSTERED)
        prim = sc.rays_primary(1920, 1080, spp=4, seed=0x5EED0001 + 2, flags=0)
        bseed = 77
    else:
        s = sc.Scene(a.triangles or 10_000_000, seed=0x5EED0001 + 3)
        prim = sc.rays_primary(1920, 1080, flags=0)
        bseed = 5
    ctx = api.Context(max_treelet_size=a.budget, device=0)
    ctx.register(s); ti = ctx.form_treelets()
    rd = torch.from_numpy(prim.view(np.uint8).reshape(-1)).to(dev)
    ctx.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(prim))
    r = ctx.device_results()
    hits = torch.as_tensor(_Dev(r.hits, len(prim) * _abi.HIT.itemsize), device=dev).cpu().numpy().view(_abi.HIT)
    rays = s.bounce(prim, hits, bseed, 0 if a.config == "C3" else 1, 0)
    rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    out = []
    for _ in range(a.reps):
        ctx.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))
        r = ctx.device_results()
        out.append({"order_ms": r.order_ms, "k1_ms": r.traverse_ms, "k3_ms": r.compact_ms, "scan_ms": r.scan_ms})
    print(json.dumps({"config": a.config, "budget": a.budget, "rays": int(len(rays)), "records_per_ray": r.n_txn / len(rays), "bytes_per_ray": r.algorithmic_bytes / len(rays),
                      "treelets": int(ti.n_treelets), "form_ms": ti.form_ms, "arena_bytes": int(s.size), "passes": out,
                      "tb_stats": ctx.tb_stats() if os.environ.get("VSRT_K1_TB", "0") != "0" else None}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

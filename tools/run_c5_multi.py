#!/usr/bin/env python
"""BASELINE.json configs[4] on N GPUs of one box: 10 M-triangle scene, 3840 x 2160 x 8 spp primary rays (66 355 200 rays),
ray ids sharded in contiguous blocks over the ranks (BVH replicated, built redundantly), every rank traces its block in
batches with device-resident rays and results, then ONE NCCL reduce of the functional counters and the treelet visit
histogram for the frame.  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_c5_multi.py

Time = per-rank sum of the library's CUDA-event batch times (K1 + scan + K3) + the reduce, max over ranks; rank 0 prints one
JSON line.  A parity property rides along: the reduced ray count and record-type histogram must equal the frame totals."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
import vsrt.api as api
from vsrt import scene as sc, _abi, shard


class _DevArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n_tri = int(os.environ.get("VSRT_C5_TRIANGLES", 10_000_000)); batch = int(os.environ.get("VSRT_C5_BATCH", 1 << 22))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ.pop("OMP_NUM_THREADS", None)
    t0 = time.time()
    s = sc.Scene(n_tri, seed=0x5EED0001 + 4)
    ctx = api.Context(max_treelet_size=512, device=local); ctx.register(s); ti = ctx.form_treelets()
    build_s = time.time() - t0
    W, H, spp = 3840, 2160, 8
    total = W * H * spp
    first, count = shard.shard_range(total, world, rank)
    # two passes over the rank's block: the first (untimed) lets the library grow its staging / output buffers to the largest
    # batch, then the counters are reset and the second pass is the measurement
    for timed in (False, True):
        if timed:
            ctx.reset_counters()
        done, ms_tot, rec = 0, 0.0, 0
        while done < count:
            m = min(batch, count - done)
            rays = sc.rays_primary(W, H, spp=spp, seed=9, first=first + done, count=m)
            rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
            ctx.trace_device(_abi.MODE_TREELET, rd.data_ptr(), m)
            r = ctx.device_results()
            ms_tot += r.traverse_ms + r.scan_ms + r.compact_ms; rec += r.n_txn; done += m
    cptr, hptr, nt = ctx.counters_device()
    csum = torch.as_tensor(_DevArray(cptr, 8 * _abi.N_SUM), device=dev).view(torch.int64).clone()
    cmax = torch.as_tensor(_DevArray(cptr + 8 * _abi.N_SUM, 8 * _abi.N_MAX), device=dev).view(torch.int64).clone()
    hist = torch.as_tensor(_DevArray(hptr, 8 * nt), device=dev).view(torch.int64).clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        shard.reduce_counters(dist, csum.clone(), cmax.clone(), hist.clone())          # warm-up of the communicator
        torch.cuda.synchronize(); dist.barrier()
    e0.record()
    if world > 1:
        shard.reduce_counters(dist, csum, cmax, hist)
    e1.record(); torch.cuda.synchronize()
    red_ms = e0.elapsed_time(e1)
    t = torch.tensor([ms_tot + red_ms, ms_tot, float(rec)], dtype=torch.float64, device=dev)
    tmax = t.clone(); tsum = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        names = list(_abi.COUNTER_FIELDS)
        cs = dict(zip(names[:_abi.N_SUM], csum.cpu().tolist())); cm = dict(zip(names[_abi.N_SUM:], cmax.cpu().tolist()))
        n_rec = int(tsum[2].item())
        assert cs["ray_count"] == total, (cs["ray_count"], total)
        assert sum(cs["mem_access_type_%d" % i] for i in range(9)) == n_rec
        assert int(hist.sum().item()) <= n_rec
        print(json.dumps({"config": "C5", "n_gpus": world, "triangles": n_tri, "rays": total, "rays_per_gpu": count, "batch": batch,
                          "ms_max_over_ranks": float(tmax[0].item()), "trace_ms_max": float(tmax[1].item()), "reduce_ms": red_ms,
                          "rays_per_s": total / float(tmax[0].item()) * 1e3, "records_per_ray": n_rec / total, "treelets": int(ti.n_treelets),
                          "treelet_form_ms": float(ti.form_ms), "scene_build_s": build_s, "hits": cs["num_hits"],
                          "max_nodes_per_ray": cm["max_nodes_per_ray"], "max_tree_depth": cm["max_tree_depth"],
                          "treelet_hist_sum": int(hist.sum().item())}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo test of the multi-GPU host logic (no GPU): ray-id sharding + counter/histogram reduce.
Each rank traces its shard with the oracle port (standing in for its GPU), the reduced counters and histogram must
equal a single-process run over the whole frame, and the concatenated per-rank traces must equal the full trace."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    from vsrt import scene as sc, _abi, shard
    import oracles
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = sc.Scene(4000, seed=12, n_blas=2, n_instances=2)          # replicated, deterministic
    W, H, spp = 40, 30, 2
    total = W * H * spp
    first, count = shard.shard_range(total, world, rank)
    rays = sc.rays_primary(W, H, spp=spp, seed=5, first=first, count=count)
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    r = orc.trace(1, rays)
    c = orc.counters()
    t = orc.tables()
    idx = np.searchsorted(t["roots"], r["treelet_ids"])
    hist = torch.from_numpy(np.bincount(idx, minlength=len(t["roots"])).astype(np.int64))
    vals = {"mem_access_type_%d" % i: c["mem_access_type_%d" % i] for i in range(9)}
    vals.update(num_hits=c["num_hits"], num_any_hits=c["num_any_hits"], n_anyhit_rays=c["n_anyhit_rays"], n_closesthit_rays=c["n_closesthit_rays"],
                tot_nodes_per_ray=c["tot_nodes_per_ray"], accessed_data_size=c["accessed_data_size"], ray_count=c["ray_count"],
                max_nodes_per_ray=c["max_nodes_per_ray"], max_tree_depth=c["max_tree_depth"])
    flat = torch.tensor([vals[k] for k in _abi.COUNTER_FIELDS], dtype=torch.int64)
    csum, cmax = flat[:_abi.N_SUM].clone(), flat[_abi.N_SUM:].clone()
    shard.reduce_counters(dist, csum, cmax, hist)
    # the scheme the CUDA library uses (csrc/reduce.cu): ONE sum-reduce of a packed header per frame -- deltas of the SUM counters,
    # every rank's MAX counters in its own pair of a world-wide table -- and u32 histogram deltas, folded into running totals.
    # Two "frames": the first half of the shard, then the rest.
    totals = np.zeros(_abi.N_SUM + _abi.N_MAX, np.int64); prev = np.zeros_like(totals); ghist = np.zeros(len(t["roots"]), np.int64); hprev = np.zeros_like(ghist)
    orc2 = oracles.PortOracle(); orc2.register(s); orc2.form(512)
    half = (count // 64) * 32
    for lo, hi in ((0, half), (half, count)):
        rr = orc2.trace(1, rays[lo:hi])
        cc = orc2.counters()
        now = np.array([{**{"mem_access_type_%d" % i: cc["mem_access_type_%d" % i] for i in range(9)}, **cc}[k] for k in _abi.COUNTER_FIELDS], dtype=np.int64)
        hdr = torch.from_numpy(shard.pack_reduce_header(now, prev, world, rank)); prev = now.copy()
        hnow = hprev + np.bincount(np.searchsorted(t["roots"], rr["treelet_ids"]), minlength=len(t["roots"]))
        dh = torch.from_numpy((hnow - hprev).astype(np.int32)); hprev = hnow
        dist.all_reduce(hdr, op=dist.ReduceOp.SUM); dist.all_reduce(dh, op=dist.ReduceOp.SUM)
        totals, overflow = shard.fold_reduce_header(totals, hdr.numpy(), world)
        assert not overflow
        ghist += dh.numpy()
    r["txns"]["address"] -= np.uint64(s.base)      # every process maps the (identical) arena at its own host address
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), first=first, count=count, offsets=r["offsets"], txns=r["txns"],
             csum=csum.numpy(), cmax=cmax.numpy(), hist=hist.numpy(), packed_totals=totals, packed_hist=ghist)
    dist.destroy_process_group()


def test_shard_ranges():
    for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
        sys.path.insert(0, p)
    from vsrt import shard
    for total in (0, 1, 31, 32, 1000, 2073600, 66355200):
        for world in (1, 2, 4, 8):
            pos = 0
            for r in range(world):
                first, count = shard.shard_range(total, world, r)
                assert first == pos and (first % 32 == 0 or count == 0)
                pos += count
            assert pos == total


def test_two_rank_reduce_matches_single_process(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    from vsrt import scene as sc, _abi
    import oracles
    s = sc.Scene(4000, seed=12, n_blas=2, n_instances=2)
    rays = sc.rays_primary(40, 30, spp=2, seed=5)
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    full = orc.trace(1, rays); c = orc.counters(); t = orc.tables()
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert parts[0]["first"] == 0 and parts[1]["first"] == parts[0]["count"] and parts[0]["count"] + parts[1]["count"] == len(rays)
    full["txns"]["address"] -= np.uint64(s.base)
    assert np.array_equal(np.concatenate([p["txns"] for p in parts]), full["txns"])      # traces concatenate in rank order
    hist = np.bincount(np.searchsorted(t["roots"], full["treelet_ids"]), minlength=len(t["roots"]))
    for p in parts:                                                                      # every rank holds the reduced values
        assert np.array_equal(p["hist"], hist)
        got = dict(zip(_abi.COUNTER_FIELDS, list(p["csum"]) + list(p["cmax"])))
        for k in _abi.COUNTER_FIELDS:
            assert got[k] == c[k], k
        # the packed single-reduce scheme over two frames arrives at the same totals (sums, maxima, histogram)
        got2 = dict(zip(_abi.COUNTER_FIELDS, list(p["packed_totals"])))
        for k in _abi.COUNTER_FIELDS:
            assert got2[k] == c[k], ("packed", k)
        assert np.array_equal(p["packed_hist"], hist)

#!/usr/bin/env python
"""bench.py -- rays/s of the functional ray-traversal hot path (traversal + access trace + treelet ids).

    python bench.py --gpus N --steps K --warmup W             # CUDA path (libvsrt.so), one process per GPU
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code (oracle/_ref), host cores

Workload (BASELINE.json configs[1]): synthetic 1M-triangle scene, 1920x1080 primary rays, full access trace +
treelet ids, traversal variant and treelet budget of the reference's shipped treelet_prefetching config
(-treelet_based_traversal 1, -max_treelet_size 512).  At N GPUs the frame gets N samples per pixel and rank r
traces the r-th contiguous ray-id block of 2,073,600 rays (weak scaling, BVH replicated, no data-path
collective); the only exchange is the per-frame NCCL all-reduce of the counters and the treelet visit histogram.
A "step" = one pass over the rank's batch: K1 traversal -> scan -> K3 trace compaction.

One JSON line on stdout (rank 0).  value = rays/s with the rays resident in HBM, timed with CUDA events on the
launching stream, max over ranks; e2e = the same through vsrt_trace_rays with pinned HOST buffers (H2D of the
rays and D2H of hits + CSR offsets + records + treelet ids inside the timed region)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 1920, 1080
N_TRI = 1_000_000
SCENE_SEED = 0x5EED0001 + 1
BUDGET = int(os.environ.get("VSRT_BENCH_BUDGET", "512"))   # configs/tested-cfgs/treelet_prefetching/gpgpusim.config:219
MODE = int(os.environ.get("VSRT_BENCH_MODE", "1"))          # -treelet_based_traversal 1 (gpgpusim.config:225)
METRIC = "rays/s (traversal + access trace + treelet ids)"
# Order of the frame's ray ids: "8x4" = pixel tiles of 8 x 4, the order in which the reference's raygen launch hands rays to
# traceRay (one-warp CTAs of 8 x 4 pixels, warp_pixel_mapping WARP_8X4, vulkan_ray_tracing.cc:3505); "0" = scanlines
_t = os.environ.get("VSRT_BENCH_TILE", "8x4").lower()
TILE = tuple(int(v) for v in _t.split("x")) if "x" in _t else None


def workload_config(n_gpus, extra=None):
    cfg = {"workload": "synthetic %dM-triangle scene, %dx%d primary rays %d spp (%d rays/GPU), traceRayWithTreelets, max_treelet_size %d B, full access trace + treelet ids"
           % (N_TRI // 1_000_000, WIDTH, HEIGHT, n_gpus, WIDTH * HEIGHT, BUDGET),
           "ray_order": ("%dx%d pixel tiles (the reference's WARP_8X4 raygen mapping, vulkan_ray_tracing.cc:3505)" % TILE) if TILE else "scanlines",
           "triangles": N_TRI, "rays_per_gpu": WIDTH * HEIGHT, "mode": "treelet", "max_treelet_size": BUDGET,
           "sharding": "rays: contiguous ray-id block per GPU; BVH replicated"}
    if extra:
        cfg.update(extra)
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """Start of the timed region: only samples written after this point are used (if there are any)."""
        try:
            self.f.flush()
            self.skip = len(open(self.f.name).read().splitlines())
        except Exception:
            self.skip = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        lines = self.f.read().splitlines()
        skip = getattr(self, "skip", 0)
        if len(lines) - skip >= 2:
            lines = lines[skip:]
        for line in lines:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def build_scene_and_rays(n_gpus, rank, flags=0):
    from vsrt import scene as sc
    t0 = time.time()
    s = sc.Scene(N_TRI, seed=SCENE_SEED)
    from vsrt import shard
    first, count = shard.shard_range(WIDTH * HEIGHT * n_gpus, n_gpus, rank)     # contiguous ray-id block of this rank
    rays = sc.rays_primary(WIDTH, HEIGHT, spp=n_gpus, seed=SCENE_SEED, flags=flags, first=first, count=count, tile=TILE)
    return s, rays, time.time() - t0


# ------------------------------------------------------------------------------------------------ reference arm
_REF = {}


def _ref_worker(args):
    lo, hi, mode = args
    orc, rays = _REF["orc"], _REF["rays"]
    r = orc.trace(mode, rays[lo:hi], cap_per_ray=256)
    return int(len(r["txns"])), int(r["txns"]["size"].sum())


def cpu_reference_setup(rays_sample):
    """Loads the CPU implementation of the path: oracle/_ref (the reference's own code) when present, else the
    C port.  Formation happens once here (the reference forms lazily on its first ray)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build_cpu()
    import oracles
    orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
    return orc


def time_cpu(orc, scene, rays, mode, cores, steps, warmup):
    """Times `steps` passes of the CPU path over `rays` with `cores` workers.  The reference's code is
    single-threaded with process-global state, so parallelism = forked worker processes that inherit the formed
    treelet maps copy-on-write; the C port uses OpenMP threads in-process."""
    import multiprocessing as mp
    t0 = time.time()
    orc.register(scene); orc.form(BUDGET)
    form_s = time.time() - t0
    n = len(rays)
    times, recs, nbytes = [], 0, 0
    last = None
    if orc.kind == "reference" and cores == 1:
        # the reference as it runs: one thread, lane after lane, in this process -- the trace is kept for the parity sample
        for it in range(warmup + steps):
            t = time.perf_counter()
            last = orc.trace(mode, rays, cap_per_ray=256)
            dt = time.perf_counter() - t
            if it >= warmup:
                times.append(dt)
            recs, nbytes = len(last["txns"]), int(last["txns"]["size"].sum())
    elif orc.kind == "reference":
        _REF["orc"], _REF["rays"] = orc, rays
        chunks = [(n * i // cores, n * (i + 1) // cores, mode) for i in range(cores)]
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            for it in range(warmup + steps):
                t = time.perf_counter()
                res = pool.map(_ref_worker, chunks)
                dt = time.perf_counter() - t
                if it >= warmup:
                    times.append(dt)
                recs, nbytes = sum(r[0] for r in res), sum(r[1] for r in res)
    else:
        for it in range(warmup + steps):
            t = time.perf_counter()
            r = orc.trace(mode, rays, cap_per_ray=256, nthreads=cores)
            dt = time.perf_counter() - t
            if it >= warmup:
                times.append(dt)
            recs, nbytes = len(r["txns"]), int(r["txns"]["size"].sum())
            last = r
    return {"rays_per_s": n * len(times) / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "form_s": form_s,
            "records_per_ray": recs / max(n, 1), "bytes_per_ray": nbytes / max(n, 1), "trace": last}


def parity_of(o, g):
    """Bitwise comparison of an oracle trace (o) with the CUDA path's (g) for the same rays: per-ray record counts, every
    {address, size, type}, every treelet id, hit flags / ids and the bit patterns of t, barycentrics and hit point."""
    oh, gh = o["hits"], g["hits"]
    checks = {
        "offsets": np.array_equal(o["offsets"], g["offsets"]),
        "txns": len(o["txns"]) == len(g["txns"]) and np.array_equal(o["txns"]["address"], g["txns"]["address"]) and
                np.array_equal(o["txns"]["size"], g["txns"]["size"]) and np.array_equal(o["txns"]["type"], g["txns"]["type"]),
        "treelet_ids": np.array_equal(o["treelet_ids"], g["treelet_ids"]),
        "hit_ids": np.array_equal(oh["hit"], gh["hit_geometry"]) and np.array_equal(oh["prim"], gh["primitive_index"]) and
                   np.array_equal(oh["geom"], gh["geometry_index"]) and np.array_equal(oh["instance_id"], gh["instance_index"]),
        "hit_t_bary_point_bits": np.array_equal(oh["t"].view(np.uint32), gh["world_min_thit"].view(np.uint32)) and
                                 np.array_equal(oh["bary"].view(np.uint32), gh["barycentric"].view(np.uint32)) and
                                 np.array_equal(oh["point"].view(np.uint32), gh["intersection_point"].view(np.uint32)),
    }
    return {"rays": int(len(o["offsets"]) - 1), "records": int(len(o["txns"])), "equal": bool(all(checks.values())),
            "checks": {k: bool(v) for k, v in checks.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vsrt import scene as sc  # noqa: F401
    scene, rays, _ = build_scene_and_rays(args.gpus, 0)
    orc = cpu_reference_setup(None)
    cores = os.cpu_count() or 1
    if orc.kind == "reference":
        cores = min(cores, 64)
    # size the per-step sample so that the whole --steps/--warmup run ends within a few minutes: calibrate on a
    # small strided sample first (the reference manages only a few thousand rays/s/core on this scene)
    cal = np.ascontiguousarray(rays[::max(1, len(rays) // 8192)][:8192])
    c = time_cpu(orc, scene, cal, MODE, cores, 1, 0)
    target_s = min(10.0, 150.0 / max(1, args.steps + args.warmup))
    sample_n = int(min(len(rays), args.ref_sample, max(16384, c["rays_per_s"] * target_s)))
    # evenly spaced over the WHOLE frame (round 1 took the first sample_n rays when the stride rounded down to 1)
    idx = (np.arange(sample_n, dtype=np.int64) * len(rays)) // sample_n
    stride = len(rays) / sample_n
    sample = np.ascontiguousarray(rays[idx])
    r = time_cpu(orc, scene, sample, MODE, cores, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u64", "data": "synthetic", "config": workload_config(args.gpus),      # the same keys as the CUDA arm's workload description
            "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": cores, "kind": orc.kind,
                             "sample": "bounded sample per step: %d rays evenly spaced over the frame (one in %.2f), %d worker %s; treelet formation %.1f s excluded"
                                       % (len(sample), stride, cores, "processes (fork, maps shared copy-on-write)" if orc.kind == "reference" else "OpenMP threads", r["form_s"])},
            "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "records_per_ray": r["records_per_ray"], "bytes_per_ray": r["bytes_per_ray"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA arm
class _DevArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from vsrt import _abi, shard
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line and nothing else (NCCL prints its version there)
    if world > 1:
        # the host-buffer calls use worker threads (record expansion / treelet ids); the ranks of one box share its cores
        os.environ.setdefault("VSRT_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // world)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import vsrt.api as api
    dev = torch.device("cuda", local)

    scene, rays, build_s = build_scene_and_rays(world, rank)
    n = len(rays)
    ctx = api.Context(max_treelet_size=BUDGET, device=local, treelet_based_traversal=MODE)
    ctx.register(scene)
    ti = ctx.form_treelets()

    stream = torch.cuda.Stream(device=dev)
    rays_pinned = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
    rays_dev = rays_pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # The path's only exchange: per-frame reduce of the counters + the treelet visit histogram, done by the library
    # (vsrt_reduce_counters: snapshot on the traversal stream, one grouped NCCL all-reduce pair on the library's own stream into
    # buffers it owns, fold into global totals), so the reduce of frame i overlaps the traversal of frame i + 1.
    if world > 1:
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    frames = {"n": 0}

    def step():
        with torch.cuda.stream(stream):
            ctx.trace_device(MODE, rays_dev.data_ptr(), n, stream.cuda_stream)
            frames["n"] += 1
            if world > 1:
                ctx.reduce_counters(stream.cuda_stream)

    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs ~100 ms to come up
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    trav_ms = scan_ms = comp_ms = 0.0
    launches = 0
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # evict the L2 between timed iterations (not timed)
            ev[i][0].record(stream)
        step()
        with torch.cuda.stream(stream):
            if i == args.steps - 1 and world > 1:
                ctx.reduce_wait(stream.cuda_stream)    # the last frame's reduce is inside the timed region
            ev[i][1].record(stream)
        r = ctx.device_results()
        trav_ms += r.traverse_ms; scan_ms += r.scan_ms; comp_ms += r.compact_ms; launches += r.kernel_launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    reduce_check = None
    if world > 1:
        # every rank holds the global totals; they must be exactly world x this rank's (every rank traces a block of the same
        # size) -- the round-1 bench printed sums that a racy reduce had doubled, so this is asserted, not just printed
        gtot, ghist = ctx.reduced()
        mine = ctx.counters()
        rec_mine = sum(mine["mem_access_type_%d" % i] for i in range(9))
        rec_glob = sum(gtot["mem_access_type_%d" % i] for i in range(9))
        rec_all = torch.tensor([rec_mine, int(ctx.treelet_histogram().sum())], dtype=torch.int64, device=dev)
        dist.all_reduce(rec_all, op=dist.ReduceOp.SUM)
        reduce_check = {"ray_count": gtot["ray_count"], "expected_ray_count": world * n * frames["n"], "records": rec_glob,
                        "expected_records": int(rec_all[0].item()), "treelet_hist_sum": int(ghist.sum()), "expected_treelet_hist_sum": int(rec_all[1].item()), "frames": frames["n"],
                        "max_nodes_per_ray": gtot["max_nodes_per_ray"], "max_tree_depth": gtot["max_tree_depth"]}
        reduce_check["reduce_ok"] = bool(reduce_check["ray_count"] == reduce_check["expected_ray_count"] and rec_glob == reduce_check["expected_records"] and
                                         reduce_check["treelet_hist_sum"] == reduce_check["expected_treelet_hist_sum"] and gtot["max_nodes_per_ray"] >= mine["max_nodes_per_ray"])
        assert reduce_check["reduce_ok"], "multi-GPU counter reduce returned wrong totals: %r" % (reduce_check,)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    res = ctx.device_results()
    n_txn, alg_bytes = res.n_txn, res.algorithmic_bytes
    t_all = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms, total_rays = float(t_all.item()), float(tot.item())
    value = total_rays * args.steps / (max_ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers (pinned): H2D rays, D2H hits + offsets + records + treelet ids
    hits_h = torch.empty(n * _abi.HIT.itemsize, dtype=torch.uint8).pin_memory()
    offs_h = torch.empty((n + 1) * 8, dtype=torch.uint8).pin_memory()
    txn_h = torch.empty(max(n_txn, 1) * 16, dtype=torch.uint8).pin_memory()
    tid_h = torch.empty(max(n_txn, 1) * 8, dtype=torch.uint8).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        return ctx.trace_into(MODE, n, rays_pinned.data_ptr(), hits_h.data_ptr(), offs_h.data_ptr(), txn_h.data_ptr(), n_txn, tid_h.data_ptr())
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert got == n_txn
    e_all = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_all, op=dist.ReduceOp.MAX)
    e2e_value = total_rays * e2e_steps / float(e_all.item())
    h2d = n * _abi.RAY.itemsize
    # what the caller receives (hits, offsets, 16-byte records, 64-bit treelet ids) and what crosses the link for it: with host
    # expansion (the library's default for this layout) a record travels as 4 packed bytes and worker threads write the 24
    delivered = n * _abi.HIT.itemsize + (n + 1) * 8 + n_txn * 16 + n_txn * 8
    host_expand = os.environ.get("VSRT_HOST_EXPAND", "1") != "0"
    d2h = n * _abi.HIT.itemsize + (n + 1) * 8 + (n_txn * 4 if host_expand else n_txn * 16)

    # ---- the same with the packed host form (vsrt_trace_rays_packed): 4-byte records + 4-byte treelet indices, expanded by the
    # caller where it consumes them (vsrt_unpack_txn).  Reported beside "e2e", which stays the full 24 bytes per record.
    rec_h = torch.empty(max(n_txn, 1) * 4, dtype=torch.uint8).pin_memory()

    # lean form: no treelet-index stream (the host derives it from vsrt_node_treelet_table, fetched once per formation), the
    # frame traced in chunks whose copies overlap the next chunk's upload + traversal
    node_table = ctx.node_treelet_table()

    def e2e_packed_step():
        return ctx.trace_packed_into(MODE, n, rays_pinned.data_ptr(), hits_h.data_ptr(), offs_h.data_ptr(), rec_h.data_ptr(), n_txn, None)
    e2e_packed_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = e2e_packed_step()
    torch.cuda.synchronize()
    e2ep_s = time.perf_counter() - t0
    assert got == n_txn
    ep_all = torch.tensor([e2ep_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ep_all, op=dist.ReduceOp.MAX)
    e2e_packed_value = total_rays * e2e_steps / float(ep_all.item())
    d2h_packed = n * _abi.HIT.itemsize + (n + 1) * 8 + n_txn * 4
    # what the lean form delivers is checked against the device-resident records of the same frame: record i expands to txns[i]
    # and node_table[record >> 3] is its treelet index
    rec_np = rec_h.numpy().view(np.uint32)[:n_txn]
    ctx.trace_device(MODE, rays_dev.data_ptr(), n)
    txn_dev, tid_dev = ctx.fetch_trace()
    tix_dev = np.searchsorted(ctx.tables()["roots"], tid_dev).astype(np.uint32)
    lean_ok = bool(np.array_equal(ctx.unpack(rec_np[:1 << 20]), txn_dev[:1 << 20]) and np.array_equal(node_table[rec_np >> 3], tix_dev))
    # ... and what the full host form delivered (the e2e call above): every record and every treelet id of the frame
    e2e_ok = bool(np.array_equal(txn_h.numpy().view(_abi.TXN)[:n_txn], txn_dev) and np.array_equal(tid_h.numpy().view(np.uint64)[:n_txn], tid_dev))
    # the PCIe link this box gives a pinned device->host copy (the floor of any host-buffer call)
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev); big_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    big_h.copy_(big); torch.cuda.synchronize()
    t0 = time.perf_counter(); big_h.copy_(big); torch.cuda.synchronize(); d2h_gbs = (1 << 30) / (time.perf_counter() - t0) / 1e9
    del big, big_h

    if rank == 0:
        peak, peak_src = peaks()
        k1_ms = trav_ms / args.steps
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        traffic = traffic_src = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_traverse_dram_bytes_per_launch")
                traffic_src = "profiles/traffic.json: ncu --set full capture %s at commit %s" % (tj.get("capture"), tj.get("commit"))
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32+u64", "data": "synthetic",
                "config": workload_config(world, {"l2": "256 MiB buffer rewritten between timed steps (L2 flush); per-step working set %.2f GB"
                                                        % ((scene.size + n * 52 + n * res_stage_bytes(ctx) + n_txn * 20) / 1e9),
                                                  "treelets": int(ti.n_treelets), "treelet_form_ms": float(ti.form_ms), "scene_build_s": build_s,
                                                  "records_per_ray": n_txn / n, "bytes_per_ray": alg_bytes / n, "arena_bytes": int(scene.size)}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                        "host_bytes_delivered_per_step": int(delivered), "matches_device_records": e2e_ok,
                        "api": "vsrt_trace_rays (host pinned buffers in, hits + CSR offsets + 16-byte records + 64-bit treelet ids in host buffers out; "
                               + ("the frame is traced in 524,288-ray windows whose records cross the link as 4 packed bytes and are expanded to the 24 bytes by the library's host threads behind the copy front)"
                                  if host_expand else "records copied as 16-byte records, ids derived on the host)")},
                "e2e_packed": {"value": e2e_packed_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_packed), "steps": e2e_steps,
                               "api": "vsrt_trace_rays_packed, lean form (host pinned buffers; hits + CSR offsets + 4-byte packed records, traced in 524,288-ray chunks with the copies overlapped; the caller expands with vsrt_unpack_txn and takes treelet ids from vsrt_node_treelet_table)",
                               "matches_device_records": lean_ok},
                "pcie": {"d2h_gbs_measured": d2h_gbs, "e2e_floor_ms": d2h / d2h_gbs / 1e6, "e2e_packed_floor_ms": d2h_packed / d2h_gbs / 1e6,
                         "note": "floor = bytes that cross the link per frame / measured pinned D2H bandwidth; the full form is bound by the host threads that write 24 bytes per record, the lean form by the link"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_traverse<TREELET>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": k1_ms,
                             "step_breakdown_ms": {"k_traverse": k1_ms, "scan": scan_ms / args.steps, "k_compact": comp_ms / args.steps}},
                }
        if world == 1 and not args.no_cpu_baseline:
            try:
                sample_n = min(n, args.cpu_sample)
                stride = max(1, n // sample_n)
                sample = np.ascontiguousarray(rays[::stride][:sample_n])
                orc = cpu_reference_setup(sample)
                r = time_cpu(orc, scene, sample, MODE, 1, 1, 0)
                line["cpu_baseline"] = {"value": r["rays_per_s"], "unit": "rays/s", "cores": 1, "kind": orc.kind,
                                        "sample": "%d rays (every %d-th ray of the frame), 1 pass, single thread as the reference runs; treelet formation %.1f s excluded"
                                                  % (len(sample), stride, r["form_s"]),
                                        "host_cores_available": os.cpu_count()}
                # the same rays through the CUDA path (host-buffer C-ABI call), compared bit for bit with what the CPU arm just
                # produced: the headline number is for THIS scene, so its parity is checked on this scene, not extrapolated
                g_tr = ctx.trace(MODE, sample)
                line["parity_sample"] = parity_of(r["trace"], g_tr)
                line["parity_sample"]["oracle"] = orc.kind
                to, tg = orc.tables(), ctx.tables()
                line["parity_sample"]["treelet_tables_equal"] = bool(all(np.array_equal(to[k], tg[k]) for k in ("roots", "counts", "node_addr", "node_size", "map_nodes", "map_roots")))
                line["parity_sample"]["equal"] = bool(line["parity_sample"]["equal"] and line["parity_sample"]["treelet_tables_equal"])
            except Exception as e:   # the baseline is reporting only; never lose the GPU line to it
                line["cpu_baseline"] = {"value": None, "unit": "rays/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        if reduce_check is not None:
            line["reduced_counters"] = reduce_check
        if world == 1 and not args.no_incoherent:
            try:
                line["incoherent"] = run_incoherent(api, torch, dev, flush, peak, args.c4_triangles)
            except Exception as e:   # reporting beside the headline; never lose the line to it
                line["incoherent"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _timed_batch(ctx, torch, dev, mode, rays, flush, reps=2):
    """One untimed pass (sizes the library's buffers), then `reps` passes with the L2 flushed before each; device time of the
    three stages from the library's CUDA events on the launching stream.  Returns (mean ms dict, device results, hits on host)."""
    from vsrt import _abi
    rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    ctx.trace_device(mode, rd.data_ptr(), len(rays))
    acc = {"order": 0.0, "k1": 0.0, "scan": 0.0, "k3": 0.0}
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        ctx.trace_device(mode, rd.data_ptr(), len(rays))
        r = ctx.device_results()
        acc["order"] += r.order_ms / reps; acc["k1"] += r.traverse_ms / reps; acc["scan"] += r.scan_ms / reps; acc["k3"] += r.compact_ms / reps
    r = ctx.device_results()
    src = torch.as_tensor(_DevArray(r.hits, len(rays) * _abi.HIT.itemsize), device=dev)
    hits = src.cpu().numpy().view(_abi.HIT)
    return acc, r, hits


def run_incoherent(api, torch, dev, flush, peak, c4_triangles):
    """BASELINE.json configs[2] and configs[3] (the incoherent-ray workloads the north star's target is stated on), one GPU,
    device-resident, same clocks as the headline: rays/s over K1 + scan + K3 and K1's fraction of the HBM roofline."""
    from vsrt import scene as sc, _abi
    out = {}
    # ---- C3: 2 M-triangle clustered scene, 1080p x 4 spp primary rays, diffuse bounces 1-4 generated from the previous hits
    s = sc.Scene(2_000_000, seed=0x5EED0001 + 2, kind=sc.CLUSTERED)
    ctx = api.Context(max_treelet_size=BUDGET, device=dev.index)
    ctx.register(s); ti = ctx.form_treelets()
    rays = sc.rays_primary(WIDTH, HEIGHT, spp=4, seed=SCENE_SEED + 1, flags=0, tile=TILE)
    tot = {"rays": 0, "ms": 0.0, "k1": 0.0, "bytes": 0, "records": 0}
    per = []
    for bounce in range(5):
        ms, r, hits = _timed_batch(ctx, torch, dev, _abi.MODE_TREELET, rays, flush)
        t = ms["order"] + ms["k1"] + ms["scan"] + ms["k3"]
        per.append({"bounce": bounce, "rays": int(len(rays)), "ms": t, "order_ms": ms["order"], "k1_ms": ms["k1"], "k3_ms": ms["k3"], "rays_per_s": len(rays) / t * 1e3,
                    "records_per_ray": r.n_txn / max(len(rays), 1), "bytes_per_ray": r.algorithmic_bytes / max(len(rays), 1)})
        if bounce > 0:
            tot["rays"] += len(rays); tot["ms"] += t; tot["k1"] += ms["k1"]; tot["bytes"] += r.algorithmic_bytes; tot["records"] += r.n_txn
        if bounce == 4:
            break
        rays = s.bounce(rays, hits, 77, bounce, 0)
        if len(rays) == 0:
            break
    ach = tot["bytes"] / max(tot["k1"], 1e-9) / 1e6
    out["C3"] = {"workload": "synthetic 2M-triangle clustered scene, 1080p x 4 spp, diffuse secondary rays of bounces 1-4 (incoherent), traceRayWithTreelets, %d B treelets" % BUDGET,
                 "value": tot["rays"] / max(tot["ms"], 1e-9) * 1e3, "unit": "rays/s", "rays": tot["rays"], "ms": tot["ms"],
                 "records_per_ray": tot["records"] / max(tot["rays"], 1), "bytes_per_ray": tot["bytes"] / max(tot["rays"], 1),
                 "treelets": int(ti.n_treelets), "treelet_form_ms": float(ti.form_ms), "arena_bytes": int(s.size), "per_bounce": per,
                 "roofline": {"bound": "hbm", "kernel": "k_traverse<TREELET>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "algorithmic_bytes": int(tot["bytes"]), "kernel_ms": tot["k1"]}}
    ctx.close(); del s
    # ---- C4 at the shipped budget: 10 M-triangle scene, bounce-1 rays of a 1080p frame
    s = sc.Scene(c4_triangles, seed=0x5EED0001 + 3)
    ctx = api.Context(max_treelet_size=BUDGET, device=dev.index)
    ctx.register(s); ti = ctx.form_treelets()
    prim = sc.rays_primary(WIDTH, HEIGHT, flags=0, tile=TILE)
    _, _, hits = _timed_batch(ctx, torch, dev, _abi.MODE_TREELET, prim, flush, reps=1)
    rays = s.bounce(prim, hits, 5, 1, 0)
    ms, r, _ = _timed_batch(ctx, torch, dev, _abi.MODE_TREELET, rays, flush)
    t = ms["order"] + ms["k1"] + ms["scan"] + ms["k3"]
    ach = r.algorithmic_bytes / max(ms["k1"], 1e-9) / 1e6
    out["C4"] = {"workload": "synthetic %dM-triangle scene, diffuse bounce-1 rays of a 1080p frame (incoherent), traceRayWithTreelets, %d B treelets" % (c4_triangles // 1_000_000, BUDGET),
                 "value": len(rays) / t * 1e3, "unit": "rays/s", "rays": int(len(rays)), "ms": t, "order_ms": ms["order"], "k1_ms": ms["k1"], "k3_ms": ms["k3"],
                 "records_per_ray": r.n_txn / len(rays), "bytes_per_ray": r.algorithmic_bytes / len(rays),
                 "treelets": int(ti.n_treelets), "treelet_form_ms": float(ti.form_ms), "arena_bytes": int(s.size),
                 "roofline": {"bound": "hbm", "kernel": "k_traverse<TREELET>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "algorithmic_bytes": int(r.algorithmic_bytes), "kernel_ms": ms["k1"]}}
    ctx.close()
    return out


def res_stage_bytes(ctx):
    return 128 * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=65536, help="rays of the frame the cpu_baseline leg traces")
    ap.add_argument("--ref-sample", type=int, default=2073600, help="rays per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incoherent", action="store_true", help="skip the C3 / C4 incoherent-ray configs reported beside the headline")
    ap.add_argument("--c4-triangles", type=int, default=10_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_cuda(args)


if __name__ == "__main__":
    main()

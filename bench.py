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


def workload_config(n_gpus, extra=None):
    cfg = {"workload": "synthetic %dM-triangle scene, %dx%d primary rays %d spp (%d rays/GPU), traceRayWithTreelets, max_treelet_size %d B, full access trace + treelet ids"
           % (N_TRI // 1_000_000, WIDTH, HEIGHT, n_gpus, WIDTH * HEIGHT, BUDGET),
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
    rays = sc.rays_primary(WIDTH, HEIGHT, spp=n_gpus, seed=SCENE_SEED, flags=flags, first=first, count=count)
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
    if orc.kind == "reference":
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
    return {"rays_per_s": n * len(times) / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "form_s": form_s,
            "records_per_ray": recs / max(n, 1), "bytes_per_ray": nbytes / max(n, 1)}


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
    stride = max(1, len(rays) // sample_n)
    sample = np.ascontiguousarray(rays[::stride][:sample_n])
    r = time_cpu(orc, scene, sample, MODE, cores, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u64", "data": "synthetic", "config": workload_config(args.gpus, {"step": "bounded sample: %d rays (every %d-th ray of the frame)" % (len(sample), stride)}),
            "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": cores, "kind": orc.kind,
                             "sample": "%d rays (every %d-th of the 1080p frame) per step, %d worker %s; treelet formation %.1f s excluded"
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
    if world > 1:
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
    cptr, hptr, nt = ctx.counters_device()
    reduce_bufs = None
    if world > 1:
        csum = torch.as_tensor(_DevArray(cptr, 8 * _abi.N_SUM), device=dev).view(torch.int64)
        cmax = torch.as_tensor(_DevArray(cptr + 8 * _abi.N_SUM, 8 * _abi.N_MAX), device=dev).view(torch.int64)
        hist = torch.as_tensor(_DevArray(hptr, 8 * nt), device=dev).view(torch.int64)
        reduce_bufs = (csum, cmax, hist)

    red_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    red_state = {"done": None, "out": None}
    # The NCCL calls are enqueued by a helper thread: their host cost (~0.25 ms per frame for three collectives) would
    # otherwise sit between two frames with the GPU idle; vsrt_trace_rays_device releases the GIL while it runs.
    red_q = red_thread = None
    if world > 1:
        import queue
        import threading
        red_q = queue.Queue()

        def _reducer():
            torch.cuda.set_device(dev)
            while True:
                item = red_q.get()
                if item is None:
                    red_q.task_done()
                    return
                ready, bufs, done = item
                with torch.cuda.stream(red_stream):
                    red_stream.wait_event(ready)
                    shard.reduce_counters(dist, *bufs)
                    done.record(red_stream)
                red_q.task_done()
        red_thread = threading.Thread(target=_reducer, daemon=True)
        red_thread.start()

    def step():
        with torch.cuda.stream(stream):
            ctx.trace_device(MODE, rays_dev.data_ptr(), n, stream.cuda_stream)
            if reduce_bufs is not None:
                # the path's only exchange: per-frame reduce of counters + treelet visit histogram, into
                # scratch copies taken on the traversal stream (the per-rank counters keep their own totals).  The
                # collectives run on a side stream, so the reduce of frame i overlaps the traversal of frame i+1.
                bufs = (reduce_bufs[0].clone(), reduce_bufs[1].clone(), reduce_bufs[2].clone())
                ready = torch.cuda.Event(); ready.record(stream)
                done = torch.cuda.Event()
                red_q.put((ready, bufs, done))
                red_state["done"] = done
                red_state["out"] = bufs
        return red_state["out"]

    def drain_reduces():
        """All queued collectives are enqueued on the side stream; the caller may now wait on red_state['done']."""
        if red_q is not None:
            red_q.join()

    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs ~100 ms to come up
    for _ in range(args.warmup):
        step()
    drain_reduces()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    trav_ms = scan_ms = comp_ms = 0.0
    launches = 0
    reduced = None
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # evict the L2 between timed iterations (not timed)
            ev[i][0].record(stream)
        reduced = step()
        with torch.cuda.stream(stream):
            if i == args.steps - 1 and red_state["done"] is not None:
                drain_reduces()
                stream.wait_event(red_state["done"])   # the last frame's reduce is inside the timed region
            ev[i][1].record(stream)
        r = ctx.device_results()
        trav_ms += r.traverse_ms; scan_ms += r.scan_ms; comp_ms += r.compact_ms; launches += r.kernel_launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if red_q is not None:
        red_q.put(None)
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
    d2h = n * _abi.HIT.itemsize + (n + 1) * 8 + n_txn * 16 + n_txn * 8

    # ---- the same with the packed host form (vsrt_trace_rays_packed): 4-byte records + 4-byte treelet indices, expanded by the
    # caller where it consumes them (vsrt_unpack_txn).  Reported beside "e2e", which stays the full 24 bytes per record.
    rec_h = torch.empty(max(n_txn, 1) * 4, dtype=torch.uint8).pin_memory()
    tix_h = torch.empty(max(n_txn, 1) * 4, dtype=torch.uint8).pin_memory()

    def e2e_packed_step():
        return ctx.trace_packed_into(MODE, n, rays_pinned.data_ptr(), hits_h.data_ptr(), offs_h.data_ptr(), rec_h.data_ptr(), n_txn, tix_h.data_ptr())
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
    d2h_packed = n * _abi.HIT.itemsize + (n + 1) * 8 + n_txn * 8

    if rank == 0:
        peak, peak_src = peaks()
        k1_ms = trav_ms / args.steps
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_traverse_dram_bytes_per_launch")
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
                        "api": "vsrt_trace_rays (host pinned buffers; hits + CSR offsets + 16-byte records + 64-bit treelet ids copied back)"},
                "e2e_packed": {"value": e2e_packed_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_packed), "steps": e2e_steps,
                               "api": "vsrt_trace_rays_packed (host pinned buffers; hits + CSR offsets + 4-byte packed records + 4-byte treelet indices; the caller expands with vsrt_unpack_txn)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_traverse<TREELET>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
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
            except Exception as e:   # the baseline is reporting only; never lose the GPU line to it
                line["cpu_baseline"] = {"value": None, "unit": "rays/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        if reduced is not None:
            line["reduced_counters"] = {"ray_count": int(reduced[0][_abi.COUNTER_FIELDS.index("ray_count")].item()),
                                        "treelet_hist_sum": int(reduced[2].sum().item())}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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

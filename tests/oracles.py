"""Test-side bindings of the two CPU oracles (TEST INFRASTRUCTURE; never imported by the product):

  RefOracle  -- oracle/_ref/libvsrt_ref.so, the reference's own function bodies compiled from /root/reference
                by oracle/build_ref.sh (process-global state: one instance at a time, reset() between scenes);
  PortOracle -- oracle/libvsrt_oracle.so, the plain-C restatement (oracle/vsrt_oracle.c).

Both expose: register(arena, delta), form(budget), tables(), trace(mode, rays) -> dict of numpy arrays.
"""
import ctypes
import os
import numpy as np
from vsrt import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvsrt_ref.so")
PORT_SO = os.path.join(ROOT, "oracle", "libvsrt_oracle.so")

OHIT = np.dtype([("hit", "<u4"), ("t", "<f4"), ("prim", "<u4"), ("geom", "<u4"), ("instance_id", "<u4"),
                 ("bary", "<f4", 3), ("point", "<f4", 3), ("n_all_hits", "<u4")])
assert OHIT.itemsize == 48
OCNT_FIELDS = (["mem_access_type_%d" % i for i in range(9)] +
               ["num_hits", "num_any_hits", "n_anyhit_rays", "n_closesthit_rays", "max_nodes_per_ray",
                "tot_nodes_per_ray", "max_tree_depth", "accessed_data_size", "ray_count"])

c_u64, c_vp = ctypes.c_uint64, ctypes.c_void_p


def have_ref():
    return os.path.exists(REF_SO)


# vsrt_prefetch_decision / ref_prefetch_decision / vo_prefetch_decision
PDEC = np.dtype([("root", np.uint64), ("votes", np.uint32), ("total", np.uint32), ("submit", np.uint32), ("n_nodes", np.uint32),
                 ("first_node", np.uint32), ("num_nodes", np.uint32)])


# one warp_intersection_table::add_intersection call (Baseline table): ref_table_event / vo_table_event
TEV = np.dtype([("table", np.uint32), ("shader_counter", np.uint32), ("hit_group_index", np.uint32), ("primitive_id", np.uint32),
                ("instance_id", np.uint32), ("tid", np.uint32), ("store_addr", np.uint64, 2), ("store_size", np.uint32, 2)])


class _Base:
    def _finish_trace(self, n, total, hits, counts, txns, tids):
        offsets = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(counts, out=offsets[1:])
        return {"hits": hits, "offsets": offsets, "txns": txns[:total], "treelet_ids": tids[:total]}


class RefOracle(_Base):
    kind = "reference"

    def __init__(self):
        L = ctypes.CDLL(REF_SO)
        L.ref_alloc_tlas.argtypes = [c_vp, c_u64, c_vp]
        L.ref_alloc_blas.argtypes = [c_vp, c_u64, c_vp]
        L.ref_form_treelets.argtypes = [c_vp]
        for f in ("ref_treelet_count", "ref_treelet_total_nodes", "ref_node_map_size", "ref_remap_size", "ref_remap_base"):
            getattr(L, f).restype = c_u64
        L.ref_treelet_table.argtypes = [c_vp] * 5
        L.ref_node_map.argtypes = [c_vp] * 2
        L.ref_remap.argtypes = [c_vp] * 2
        L.ref_trace.restype = ctypes.c_int64
        L.ref_trace.argtypes = [c_vp, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, ctypes.c_int]
        L.ref_get_counters.argtypes = [c_vp]
        L.ref_config.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int]
        L.ref_sort_trace.argtypes = [ctypes.c_int, c_u64, c_vp, c_vp]
        L.ref_prefetch_vote.restype = ctypes.c_int64
        L.ref_prefetch_vote.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_u64]
        L.ref_set_treelet_metadata.argtypes = [c_u64, ctypes.c_uint]
        L.ref_trace_tables.restype = ctypes.c_int64
        L.ref_trace_tables.argtypes = [c_vp, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_u64]
        L.ref_table_bases.argtypes = [c_vp, c_vp]
        L.ref_trace_coalescing.restype = ctypes.c_int64
        L.ref_trace_coalescing.argtypes = [c_vp, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_u64, c_vp, c_vp, c_u64, c_vp]
        L.ref_schedule_pick.restype = ctypes.c_int64
        L.ref_schedule_pick.argtypes = [ctypes.c_int, c_u64, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp]
        self.L = L
        self.tlas = None

    def register(self, arena, delta=0, blas_delta=None, remap=False, stride=0):
        """delta: device address - host address for the TLAS (and every BLAS unless blas_delta is given)."""
        self.L.ref_reset()
        self._remap, self._stride = remap, stride
        self.arena, self.delta = arena, delta
        self.L.ref_alloc_tlas(arena.tlas, arena.size - arena.tlas_offset, arena.tlas + delta)
        for i, (off, size) in enumerate(arena.blas):
            d = delta if blas_delta is None else blas_delta[i]
            self.L.ref_alloc_blas(arena.base + off, size, arena.base + off + d)
        self.tlas = arena.tlas

    def form(self, budget):
        self.L.ref_config(budget, 1 if self._remap else 0, self._stride, 0)
        self.L.ref_form_treelets(self.tlas)

    def tables(self):
        L = self.L
        nt, nn, nm = L.ref_treelet_count(), L.ref_treelet_total_nodes(), L.ref_node_map_size()
        roots = np.zeros(nt, np.uint64); counts = np.zeros(nt, np.uint32); meta = np.zeros(nt, np.uint32)
        na = np.zeros(nn, np.uint64); ns = np.zeros(nn, np.uint32)
        L.ref_treelet_table(_abi.ptr(roots), _abi.ptr(counts), _abi.ptr(meta), _abi.ptr(na), _abi.ptr(ns))
        mk = np.zeros(nm, np.uint64); mv = np.zeros(nm, np.uint64)
        L.ref_node_map(_abi.ptr(mk), _abi.ptr(mv))
        return {"roots": roots, "counts": counts, "meta_idx": meta, "node_addr": na, "node_size": ns,
                "map_nodes": mk, "map_roots": mv}

    def remap_table(self):
        n = self.L.ref_remap_size()
        o = np.zeros(n, np.uint64); m = np.zeros(n, np.uint64)
        self.L.ref_remap(_abi.ptr(o), _abi.ptr(m))
        return self.L.ref_remap_base(), o, m

    def trace(self, mode, rays, cap_per_ray=512, keep_stdout=False):
        n = len(rays)
        hits = np.zeros(n, OHIT); counts = np.zeros(n, np.uint32)
        cap = max(1024, n * cap_per_ray)
        txns = np.zeros(cap, _abi.TXN); tids = np.zeros(cap, np.uint64)
        total = self.L.ref_trace(self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(counts), _abi.ptr(txns), cap,
                                 _abi.ptr(tids), 1 if keep_stdout else 0)
        if total < 0:
            return self.trace(mode, rays, cap_per_ray * 4)   # NB: re-runs the rays (counters double count)
        return self._finish_trace(n, total, hits, counts, txns, tids)

    def sort_trace(self, method, trace):
        """rt_unit::sort_mem_accesses (shader.cc:3012) on every ray of a trace dict; returns the sorted record array."""
        t = trace["txns"].copy()
        self.L.ref_sort_trace(method, len(trace["offsets"]) - 1, _abi.ptr(trace["offsets"]), _abi.ptr(t))
        return t

    def dump_as(self, out_dir, arena, desc_size):
        """The reference's own dump_descriptor_set_for_AS(split_files = true) for `arena` (BLAS headers handed over as the driver
        does through gpgpusim_pass_child_addr): files <out_dir>/gpgpusimShaders/0_0.as{main,back,front,metadata}."""
        os.makedirs(os.path.join(out_dir, "gpgpusimShaders"), exist_ok=True)
        kids = (c_vp * len(arena.blas))(*[arena.base + off for off, _ in arena.blas])
        self.L.ref_dump_as.argtypes = [ctypes.c_char_p, c_vp, ctypes.c_uint32, c_vp, ctypes.c_uint32]
        self.L.ref_dump_as((out_dir.rstrip("/") + "/").encode(), arena.tlas, desc_size, kids, len(arena.blas))
        return os.path.join(out_dir, "gpgpusimShaders", "0_0")

    def table_bases(self):
        a, b = c_u64(), c_u64()
        self.L.ref_table_bases(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def table_events(self, mode, rays):
        """The reference's own Baseline intersection / any-hit tables while tracing `rays` as CTA rows of 32 threads: per
        ray the add_intersection calls (rows, values, the two stores) and the any-hit Hit_data."""
        n = len(rays)
        counts = np.zeros(n, np.uint32)
        total = self.L.ref_trace_tables(self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(counts), None, None, 0)
        ev = np.zeros(total, TEV); ah = np.zeros(total, OHIT)
        self.L.ref_trace_tables(self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(counts), _abi.ptr(ev), _abi.ptr(ah), total)
        return counts, ev, ah

    def trace_coalescing(self, mode, rays):
        """The reference traversal with its own Function_Call_Coalescing intersection table in the loop (a fresh table per
        32 rays): per-ray transaction lists with the Intersection_Table_Load records merged in, and the store lists.  Table
        addresses come back as (offset from the table base) | 1 << 63.  Returns (txn_offsets, txns, store_offsets, stores,
        sizeof(Coalescing_Entry))."""
        n = len(rays)
        to = np.zeros(n + 1, np.uint64); so = np.zeros(n + 1, np.uint64); es = ctypes.c_uint32()
        self.L.ref_trace_coalescing(self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(to), None, 0, _abi.ptr(so), None, 0, ctypes.byref(es))
        tx = np.zeros(int(to[n]), _abi.TXN); st = np.zeros(int(so[n]), _abi.STORE)
        self.L.ref_trace_coalescing(self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(to), _abi.ptr(tx), len(tx), _abi.ptr(so), _abi.ptr(st), len(st), ctypes.byref(es))
        return to, tx, so, st, es.value

    def schedule_pick(self, trace, scheduler, last_prefetched, warp_ray_ids, stalled=None, front=None):
        """rt_unit::schedule_next_warp (shader.cc:4307-4392) for one unit."""
        ids = np.ascontiguousarray(warp_ray_ids, np.uint64)
        st = None if stalled is None else np.ascontiguousarray(stalled, np.uint8)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        return self.L.ref_schedule_pick(scheduler, int(last_prefetched), len(ids) // 32, _abi.ptr(ids), _abi.ptr(st), _abi.ptr(trace["offsets"]),
                                        _abi.ptr(fr), _abi.ptr(trace["txns"]))

    def prefetch_vote(self, trace, ray_ids, heuristic, threshold=0.0, front=None, metadata=None):
        """The treelet-prefetch vote block of rt_unit::cycle (shader.cc:3419-3685) for one group of rays."""
        ids = np.ascontiguousarray(ray_ids, np.uint64)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        if metadata:
            self.L.ref_set_treelet_metadata(metadata[0], metadata[1])
        dec = np.zeros(1, PDEC)
        args = (heuristic, threshold, 1 if metadata else 0, len(ids), _abi.ptr(ids), _abi.ptr(trace["offsets"]),
                _abi.ptr(fr) if fr is not None else None, _abi.ptr(trace["txns"]), _abi.ptr(dec))
        n = self.L.ref_prefetch_vote(*args, None, None, 0)
        ca = np.zeros(n, np.uint64); co = np.zeros(n, np.uint64)
        self.L.ref_prefetch_vote(*args, _abi.ptr(ca), _abi.ptr(co), n)
        return dec[0], ca, co

    def counters(self):
        a = np.zeros(len(OCNT_FIELDS), np.uint64)
        self.L.ref_get_counters(_abi.ptr(a))
        return dict(zip(OCNT_FIELDS, (int(x) for x in a)))


class PortOracle(_Base):
    kind = "port"

    def __init__(self):
        L = ctypes.CDLL(PORT_SO)
        L.vo_create.restype = c_vp
        L.vo_destroy.argtypes = [c_vp]
        L.vo_alloc_tlas.argtypes = [c_vp, c_vp, c_u64, c_u64]
        L.vo_alloc_blas.argtypes = [c_vp, c_vp, c_u64, c_u64]
        L.vo_form_treelets.argtypes = [c_vp, c_vp, ctypes.c_int]
        for f in ("vo_treelet_count", "vo_treelet_total_nodes", "vo_node_map_size", "vo_total_bvh_size"):
            getattr(L, f).restype = c_u64; getattr(L, f).argtypes = [c_vp]
        L.vo_treelet_table.argtypes = [c_vp] * 6
        L.vo_node_map.argtypes = [c_vp] * 3
        L.vo_treelet_remap.restype = c_u64
        L.vo_treelet_remap.argtypes = [c_vp, c_u64, ctypes.c_uint32, c_vp, c_vp]
        L.vo_trace.restype = ctypes.c_int64
        L.vo_trace.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, ctypes.c_int]
        L.vo_get_counters.argtypes = [c_vp, c_vp]
        L.vo_reset_counters.argtypes = [c_vp]
        L.vo_sort_trace.argtypes = [c_vp, ctypes.c_int, c_u64, c_vp, c_vp]
        L.vo_table_events.restype = ctypes.c_int64
        L.vo_table_events.argtypes = [c_vp, c_vp, ctypes.c_int, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp, c_u64, c_u64, c_vp, c_vp, c_vp, c_u64]
        L.vo_set_proc_sink.argtypes = [c_vp, c_vp, c_u64]
        L.vo_coalescing_events.argtypes = [c_u64, c_vp, c_vp, c_vp]
        L.vo_proc_sink_count.restype = c_u64
        L.vo_proc_sink_count.argtypes = [c_vp]
        L.vo_schedule_pick.restype = ctypes.c_int64
        L.vo_schedule_pick.argtypes = [c_vp, ctypes.c_int, c_u64, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.vo_prefetch_vote.restype = ctypes.c_int64
        L.vo_prefetch_vote.argtypes = [c_vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, c_u64, ctypes.c_uint32, c_u64, c_vp, c_vp, c_vp, c_vp,
                                       c_vp, c_vp, c_vp, c_u64]
        self.L = L
        self.h = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vo_destroy(self.h); self.h = None

    def register(self, arena, delta=0, blas_delta=None, remap=False, stride=0):
        if self.h:
            self.L.vo_destroy(self.h)
        self.h = self.L.vo_create()
        self.arena, self.delta = arena, delta
        self.L.vo_alloc_tlas(self.h, arena.tlas, arena.size - arena.tlas_offset, arena.tlas + delta)
        for i, (off, size) in enumerate(arena.blas):
            d = delta if blas_delta is None else blas_delta[i]
            self.L.vo_alloc_blas(self.h, arena.base + off, size, arena.base + off + d)
        self.tlas = arena.tlas

    def form(self, budget):
        rc = self.L.vo_form_treelets(self.h, self.tlas, budget)
        if rc != 0:
            raise RuntimeError("vo_form_treelets: %d" % rc)

    def tables(self):
        L, h = self.L, self.h
        nt, nn, nm = L.vo_treelet_count(h), L.vo_treelet_total_nodes(h), L.vo_node_map_size(h)
        roots = np.zeros(nt, np.uint64); counts = np.zeros(nt, np.uint32); meta = np.zeros(nt, np.uint32)
        na = np.zeros(nn, np.uint64); ns = np.zeros(nn, np.uint32)
        L.vo_treelet_table(h, _abi.ptr(roots), _abi.ptr(counts), _abi.ptr(meta), _abi.ptr(na), _abi.ptr(ns))
        mk = np.zeros(nm, np.uint64); mv = np.zeros(nm, np.uint64)
        L.vo_node_map(h, _abi.ptr(mk), _abi.ptr(mv))
        return {"roots": roots, "counts": counts, "meta_idx": meta, "node_addr": na, "node_size": ns,
                "map_nodes": mk, "map_roots": mv}

    def remap_table(self, base, stride):
        n = self.L.vo_treelet_remap(self.h, base, stride, None, None)
        o = np.zeros(n, np.uint64); m = np.zeros(n, np.uint64)
        self.L.vo_treelet_remap(self.h, base, stride, _abi.ptr(o), _abi.ptr(m))
        order = np.argsort(o, kind="stable")
        return o[order], m[order]

    def trace(self, mode, rays, cap_per_ray=512, nthreads=1, want_trace=True):
        n = len(rays)
        hits = np.zeros(n, OHIT); counts = np.zeros(n, np.uint32)
        if not want_trace:
            total = self.L.vo_trace(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(counts), None, 0, None, nthreads)
            return {"hits": hits, "counts": counts, "total": total}
        cap = max(1024, n * cap_per_ray)
        txns = np.zeros(cap, _abi.TXN); tids = np.zeros(cap, np.uint64)
        total = self.L.vo_trace(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(counts), _abi.ptr(txns), cap,
                                _abi.ptr(tids), nthreads)
        if total < -(1 << 62):
            raise RuntimeError("vo_trace error %d" % (-(total + (1 << 63))))
        if total < 0:
            return self.trace(mode, rays, cap_per_ray * 4, nthreads)
        return self._finish_trace(n, total, hits, counts, txns, tids)

    def sort_trace(self, method, trace):
        t = trace["txns"].copy()
        self.L.vo_sort_trace(self.h, method, len(trace["offsets"]) - 1, _abi.ptr(trace["offsets"]), _abi.ptr(t))
        return t

    def prefetch_vote(self, trace, ray_ids, heuristic, threshold=0.0, front=None, metadata=None):
        ids = np.ascontiguousarray(ray_ids, np.uint64)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        mb, mp = metadata if metadata else (0, 0)
        dec = np.zeros(1, PDEC)
        args = (self.h, heuristic, threshold, 1 if metadata else 0, mb, mp, len(ids), _abi.ptr(ids), _abi.ptr(trace["offsets"]),
                _abi.ptr(fr) if fr is not None else None, _abi.ptr(trace["txns"]), _abi.ptr(dec))
        n = self.L.vo_prefetch_vote(*args, None, None, 0)
        ca = np.zeros(n, np.uint64); co = np.zeros(n, np.uint64)
        self.L.vo_prefetch_vote(*args, _abi.ptr(ca), _abi.ptr(co), n)
        return dec[0], ca, co

    def schedule_pick(self, trace, scheduler, last_prefetched, warp_ray_ids, stalled=None, front=None):
        ids = np.ascontiguousarray(warp_ray_ids, np.uint64)
        st = None if stalled is None else np.ascontiguousarray(stalled, np.uint8)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        return self.L.vo_schedule_pick(self.h, scheduler, int(last_prefetched), len(ids) // 32, _abi.ptr(ids), _abi.ptr(st), _abi.ptr(trace["offsets"]),
                                       _abi.ptr(fr), _abi.ptr(trace["txns"]))

    def table_events(self, mode, rays, itab_base, atab_base, tid_x=None):
        sink = np.zeros(max(16, 64 * len(rays)), np.uint64)
        self.L.vo_set_proc_sink(self.h, _abi.ptr(sink), len(sink))
        t = self.trace(mode, rays)
        assert self.L.vo_proc_sink_count(self.h) <= len(sink)
        self.L.vo_set_proc_sink(self.h, None, 0)
        n = len(rays)
        counts = np.zeros(n, np.uint32)
        tx = None if tid_x is None else np.ascontiguousarray(tid_x, np.uint8)
        args = (self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(t["offsets"]), _abi.ptr(t["txns"]), _abi.ptr(tx), _abi.ptr(sink), itab_base, atab_base, _abi.ptr(counts))
        total = self.L.vo_table_events(*args, None, None, 0)
        ev = np.zeros(total, TEV); ah = np.zeros(total, OHIT)
        self.L.vo_table_events(*args, _abi.ptr(ev), _abi.ptr(ah), total)
        return counts, ev, ah

    def coalescing_events(self, event_offsets, events):
        """Coalescing_warp_intersection_table::add_intersection (intersection_table.cc:43-98) replayed over table events
        (port layout: vo_table_event)."""
        offs = np.ascontiguousarray(event_offsets, np.uint64); ev = np.ascontiguousarray(events, TEV)
        out = np.zeros(len(ev), _abi.CEV)
        assert self.L.vo_coalescing_events(len(offs) - 1, _abi.ptr(offs), _abi.ptr(ev), _abi.ptr(out)) == 0
        return out

    def trace_remapped(self, mode, rays, base, stride, budget):
        """-remap_to_treelet_layout 1 (vulkan_ray_tracing.cc:1682,:1763,...): the same visit sequence with every record
        address sent through original_bvh_to_treelet_bvh_mapping and every treelet id replaced by the remapped root
        (treelet i of the ascending root order sits at base + i * (max_treelet_size + stride))."""
        t = self.trace(mode, rays)
        o, m = self.remap_table(base, stride)
        idx = np.searchsorted(o, t["txns"]["address"])
        assert np.array_equal(o[idx], t["txns"]["address"]), "a traced address is missing from the remap table"
        txns = t["txns"].copy(); txns["address"] = m[idx]
        roots = self.tables()["roots"]
        ridx = np.searchsorted(roots, t["treelet_ids"])
        assert np.array_equal(roots[ridx], t["treelet_ids"])
        tids = (np.uint64(base) + ridx.astype(np.uint64) * np.uint64(budget + stride)).astype(np.uint64)
        return {"hits": t["hits"], "offsets": t["offsets"], "txns": txns, "treelet_ids": tids}

    def counters(self):
        a = np.zeros(len(OCNT_FIELDS), np.uint64)
        self.L.vo_get_counters(self.h, _abi.ptr(a))
        return dict(zip(OCNT_FIELDS, (int(x) for x in a)))

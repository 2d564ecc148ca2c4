"""ctypes binding of libvsrt.so (include/vsrt.h).  Mirrors the reference's call surface for this path:

    reference (VulkanRayTracing statics)            here
    ---------------------------------------------   -----------------------------------------
    gpgpusim_allocTLAS / allocBLAS                  Context.alloc_tlas / alloc_blas / register(arena)
    createTreelets (lazy, first ray)                Context.form_treelets
    addrToTreeletID / isTreeletRoot                 Context.addr_to_treelet / is_treelet_root
    traceRay / traceRayWithTreelets (per lane)      Context.trace(mode, rays)  (a batch of lanes)
    g_rt_* counters                                 Context.counters()

There is no Python/CPU implementation behind this module: if libvsrt.so is missing or no CUDA device is
usable, load()/Context() raise."""
import ctypes
import os
import numpy as np
from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# VSRT_LIB selects another build of the same library (A/B variants compiled beside it; tools/ab_variants.sh)
LIB_PATH = os.environ.get("VSRT_LIB") or os.path.join(os.path.dirname(_HERE), "libvsrt.so")
_lib = None

c_u64, c_u32, c_vp, c_int = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int

# every symbol include/vsrt.h declares (tests check the library exports all of them)
SYMBOLS = ["vsrt_default_config", "vsrt_create", "vsrt_destroy", "vsrt_last_error", "vsrt_config_parse",
           "vsrt_alloc_tlas", "vsrt_alloc_blas", "vsrt_commit", "vsrt_form_treelets", "vsrt_treelet_info_get",
           "vsrt_treelet_table", "vsrt_node_map", "vsrt_treelet_remap", "vsrt_set_treelet_layout_base", "vsrt_addr_to_treelet", "vsrt_is_treelet_root",
           "vsrt_treelet_metadata_idx", "vsrt_trace_rays", "vsrt_trace_fetch", "vsrt_trace_ray_warp",
           "vsrt_trace_rays_device", "vsrt_trace_device_results", "vsrt_get_counters", "vsrt_reset_counters",
           "vsrt_counters_device", "vsrt_get_treelet_histogram", "vsrt_sort_trace", "vsrt_prefetch_vote", "vsrt_prefetch_chunks", "vsrt_schedule_pick",
           "vsrt_table_events", "vsrt_table_event_stores", "vsrt_coalescing_events", "vsrt_coalescing_event_stores", "vsrt_coalescing_event_load",
           "vsrt_packed_layout_get", "vsrt_trace_fetch_packed", "vsrt_trace_rays_packed", "vsrt_unpack_txns",
           "vsrt_as_dump_write", "vsrt_as_dump_read", "vsrt_as_dump_free", "vsrt_register_as_image",
           "vsrt_comm_unique_id", "vsrt_comm_init", "vsrt_comm_attach", "vsrt_comm_destroy", "vsrt_reduce_counters", "vsrt_reduce_wait",
           "vsrt_reduced_get", "vsrt_reduced_device", "vsrt_node_treelet_table",
           "vsrt_enable_node_histogram", "vsrt_get_node_histogram", "vsrt_reduced_get_node_histogram"]


class VsrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vsrt error %d (%s): %s" % (code, _abi.ERRORS.get(code, "?"), msg))
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libvsrt.so is not built (%s); run __graft_entry__.build() -- there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for s in SYMBOLS:
        getattr(L, s)
    L.vsrt_default_config.argtypes = [ctypes.POINTER(_abi.Config)]
    L.vsrt_create.argtypes = [ctypes.POINTER(_abi.Config), ctypes.POINTER(c_vp)]
    L.vsrt_destroy.argtypes = [c_vp]
    L.vsrt_last_error.restype = ctypes.c_char_p
    L.vsrt_last_error.argtypes = [c_vp]
    L.vsrt_config_parse.argtypes = [ctypes.POINTER(_abi.Config), ctypes.c_char_p]
    L.vsrt_alloc_tlas.argtypes = [c_vp, c_vp, c_u64, c_u64]
    L.vsrt_alloc_blas.argtypes = [c_vp, c_vp, c_u64, c_u64]
    L.vsrt_commit.argtypes = [c_vp]
    L.vsrt_form_treelets.argtypes = [c_vp, c_vp, c_u32]
    L.vsrt_treelet_info_get.argtypes = [c_vp, ctypes.POINTER(_abi.TreeletInfo)]
    L.vsrt_treelet_table.argtypes = [c_vp] * 5
    L.vsrt_node_map.argtypes = [c_vp] * 3
    L.vsrt_treelet_remap.argtypes = [c_vp, c_u64, ctypes.POINTER(c_u64), c_vp, c_vp]
    L.vsrt_set_treelet_layout_base.argtypes = [c_vp, c_u64]
    L.vsrt_addr_to_treelet.argtypes = [c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_is_treelet_root.argtypes = [c_vp, c_u64]
    L.vsrt_treelet_metadata_idx.argtypes = [c_vp, c_u64, ctypes.POINTER(c_u32)]
    L.vsrt_trace_rays.argtypes = [c_vp, c_vp, c_int, c_u64, c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, ctypes.POINTER(c_u64)]
    L.vsrt_trace_fetch.argtypes = [c_vp, c_vp, c_u64, c_vp]
    L.vsrt_sort_trace.argtypes = [c_vp, c_int]
    L.vsrt_table_events.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_table_event_stores.argtypes = [c_vp, c_u64, c_vp]
    L.vsrt_table_event_stores.restype = None
    L.vsrt_coalescing_events.argtypes = [c_vp, c_u64, c_vp, c_vp, c_vp]
    L.vsrt_packed_layout_get.argtypes = [c_vp, c_vp, ctypes.POINTER(_abi.PackedLayout)]
    L.vsrt_trace_fetch_packed.argtypes = [c_vp, c_vp, c_u64, c_vp]
    L.vsrt_trace_rays_packed.argtypes = [c_vp, c_vp, c_int, c_u64, c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, ctypes.POINTER(c_u64)]
    L.vsrt_unpack_txns.argtypes = [ctypes.POINTER(_abi.PackedLayout), c_vp, c_u64, c_vp]
    L.vsrt_unpack_txns.restype = None
    L.vsrt_coalescing_event_stores.argtypes = [c_vp, c_vp, c_u64, c_vp]
    L.vsrt_coalescing_event_stores.restype = c_u32
    L.vsrt_coalescing_event_load.argtypes = [c_u32, c_u64, c_vp]
    L.vsrt_coalescing_event_load.restype = None
    L.vsrt_as_dump_write.argtypes = [ctypes.c_char_p, c_vp, c_u64, c_vp, c_u32, c_u64, c_u64]
    L.vsrt_as_dump_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(c_vp), ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]
    L.vsrt_as_dump_free.argtypes = [c_vp]
    L.vsrt_register_as_image.argtypes = [c_vp, c_vp, c_u64, c_u64, ctypes.c_int64, ctypes.POINTER(c_u32)]
    L.vsrt_schedule_pick.argtypes = [c_vp, c_int, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    L.vsrt_prefetch_vote.argtypes = [c_vp, ctypes.POINTER(_abi.PrefetchConfig), c_u64, c_vp, c_vp, c_vp, c_vp]
    L.vsrt_prefetch_chunks.argtypes = [c_vp, ctypes.POINTER(_abi.PrefetchConfig), c_u64, c_vp, c_vp, c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_trace_ray_warp.argtypes = [c_vp, c_vp, c_u32, c_vp, c_vp, c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_trace_rays_device.argtypes = [c_vp, c_vp, c_int, c_u64, c_vp, c_vp, ctypes.POINTER(c_u64)]
    L.vsrt_trace_device_results.argtypes = [c_vp, ctypes.POINTER(_abi.DeviceResults)]
    L.vsrt_get_counters.argtypes = [c_vp, c_vp]
    L.vsrt_reset_counters.argtypes = [c_vp]
    L.vsrt_counters_device.argtypes = [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_u64)]
    L.vsrt_get_treelet_histogram.argtypes = [c_vp, c_vp, c_u64]
    L.vsrt_node_treelet_table.argtypes = [c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_enable_node_histogram.argtypes = [c_vp, c_int]
    L.vsrt_get_node_histogram.argtypes = [c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]
    L.vsrt_reduced_get_node_histogram.argtypes = [c_vp, c_vp, c_u64]
    L.vsrt_comm_unique_id.argtypes = [c_vp]
    L.vsrt_comm_init.argtypes = [c_vp, c_u32, c_u32, c_vp]
    L.vsrt_comm_attach.argtypes = [c_vp, c_vp, c_u32, c_u32]
    L.vsrt_comm_destroy.argtypes = [c_vp]
    L.vsrt_reduce_counters.argtypes = [c_vp, c_vp]
    L.vsrt_reduce_wait.argtypes = [c_vp, c_vp]
    L.vsrt_reduced_get.argtypes = [c_vp, c_vp, c_vp, c_u64]
    L.vsrt_reduced_device.argtypes = [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_u64)]
    _lib = L
    return L


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0); hand the 128 bytes to the other ranks by any means."""
    L = load()
    b = (ctypes.c_uint8 * _abi.COMM_ID_BYTES)()
    rc = L.vsrt_comm_unique_id(b)
    if rc:
        raise VsrtError(rc, L.vsrt_last_error(None).decode())
    return bytes(b)


def parse_config(text):
    L = load()
    cfg = _abi.Config()
    L.vsrt_default_config(ctypes.byref(cfg))
    L.vsrt_config_parse(ctypes.byref(cfg), text.encode())
    return cfg


def write_as_dump(prefix, arena, desc_size=None, back_buffer=0, front_buffer=20 * 1024):
    """dump_AS files for a host arena (TLAS at arena.tlas, BLAS headers at arena.blas offsets)."""
    L = load()
    kids = (c_vp * len(arena.blas))(*[arena.base + off for off, _ in arena.blas])
    if desc_size is None:   # the TLAS buffer: up to the first BLAS above it, or the end of the arena
        above = sorted(off for off, _ in arena.blas if off > arena.tlas_offset)
        desc_size = (above[0] if above else arena.size) - arena.tlas_offset
    rc = L.vsrt_as_dump_write(prefix.encode(), arena.tlas, desc_size, kids, len(arena.blas), back_buffer, front_buffer)
    if rc:
        raise VsrtError(rc, "vsrt_as_dump_write(%s)" % prefix)


class AsImage:
    """Host image rebuilt from <prefix>.asmain/.asback/.asfront/.asmetadata."""

    def __init__(self, prefix):
        self.L = load()
        p, n, t = c_vp(), c_u64(), c_u64()
        rc = self.L.vsrt_as_dump_read(prefix.encode(), ctypes.byref(p), ctypes.byref(n), ctypes.byref(t))
        if rc:
            raise VsrtError(rc, "vsrt_as_dump_read(%s)" % prefix)
        self.ptr, self.size, self.tlas_offset = p.value, n.value, t.value
        self.base = self.ptr
        self.bytes = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(self.size,))

    @property
    def tlas(self):
        return self.ptr + self.tlas_offset

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.vsrt_as_dump_free(self.ptr); self.ptr = None


class Context:
    def __init__(self, max_treelet_size=49152, device=-1, treelet_based_traversal=1, remap_to_treelet_layout=0,
                 treelet_remap_stride=0, stack_entries=96, ray_order=0, config=None):
        L = load()
        self.L = L
        cfg = _abi.Config()
        L.vsrt_default_config(ctypes.byref(cfg))
        if config is not None:
            cfg = config
        else:
            cfg.device = device
            cfg.max_treelet_size = max_treelet_size
            cfg.treelet_based_traversal = treelet_based_traversal
            cfg.remap_to_treelet_layout = remap_to_treelet_layout
            cfg.treelet_remap_stride = treelet_remap_stride
            cfg.stack_entries = stack_entries
            cfg.ray_order = ray_order
        self.cfg = cfg
        h = c_vp()
        rc = L.vsrt_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise VsrtError(rc, L.vsrt_last_error(None).decode())
        self.h = h
        self.tlas = None
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.vsrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise VsrtError(rc, self.L.vsrt_last_error(self.h).decode())
        return rc

    # ---- registration -------------------------------------------------------------------------------------
    def alloc_tlas(self, host_addr, size, dev_addr):
        self._ck(self.L.vsrt_alloc_tlas(self.h, host_addr, size, dev_addr))
        self.tlas = host_addr

    def alloc_blas(self, host_addr, size, dev_addr):
        self._ck(self.L.vsrt_alloc_blas(self.h, host_addr, size, dev_addr))

    def register(self, arena, delta=0, blas_delta=None):
        """Register a scene.Arena the way Mesa would: one allocTLAS + one allocBLAS per BLAS header."""
        self._keep.append(arena)
        self.alloc_tlas(arena.tlas, arena.size - arena.tlas_offset, arena.tlas + delta)
        for i, (off, size) in enumerate(arena.blas):
            d = delta if blas_delta is None else blas_delta[i]
            self.alloc_blas(arena.base + off, size, arena.base + off + d)

    def register_image(self, image, delta=0):
        """Registers the TLAS and the BLASes of an AsImage; returns the number of BLASes found."""
        n = c_u32()
        self._ck(self.L.vsrt_register_as_image(self.h, image.ptr, image.size, image.tlas_offset, delta, ctypes.byref(n)))
        self.tlas = image.tlas
        self._image = image
        return n.value

    def commit(self):
        self._ck(self.L.vsrt_commit(self.h))

    # ---- treelets -----------------------------------------------------------------------------------------
    def form_treelets(self, budget=0):
        self._ck(self.L.vsrt_form_treelets(self.h, self.tlas, budget))
        return self.treelet_info()

    def treelet_info(self):
        ti = _abi.TreeletInfo()
        self._ck(self.L.vsrt_treelet_info_get(self.h, ctypes.byref(ti)))
        return ti

    def tables(self):
        ti = self.treelet_info()
        roots = np.zeros(ti.n_treelets, np.uint64); offs = np.zeros(ti.n_treelets + 1, np.uint64)
        na = np.zeros(ti.n_list_entries, np.uint64); ns = np.zeros(ti.n_list_entries, np.uint32)
        self._ck(self.L.vsrt_treelet_table(self.h, _abi.ptr(roots), _abi.ptr(offs), _abi.ptr(na), _abi.ptr(ns)))
        mk = np.zeros(ti.n_mapped_nodes, np.uint64); mv = np.zeros(ti.n_mapped_nodes, np.uint64)
        self._ck(self.L.vsrt_node_map(self.h, _abi.ptr(mk), _abi.ptr(mv)))
        return {"roots": roots, "counts": np.diff(offs).astype(np.uint32), "meta_idx": np.arange(ti.n_treelets, dtype=np.uint32),
                "node_addr": na, "node_size": ns, "map_nodes": mk, "map_roots": mv, "offsets": offs}

    def remap_table(self, base):
        n = c_u64()
        self._ck(self.L.vsrt_treelet_remap(self.h, base, ctypes.byref(n), None, None))
        o = np.zeros(n.value, np.uint64); m = np.zeros(n.value, np.uint64)
        self._ck(self.L.vsrt_treelet_remap(self.h, base, ctypes.byref(n), _abi.ptr(o), _abi.ptr(m)))
        return o, m

    def set_treelet_layout_base(self, base):
        """-remap_to_treelet_layout: where the reference's gpgpusim_malloc placed treelet_layout_bvh."""
        self._ck(self.L.vsrt_set_treelet_layout_base(self.h, int(base)))

    def addr_to_treelet(self, addr):
        r = c_u64()
        self._ck(self.L.vsrt_addr_to_treelet(self.h, addr, ctypes.byref(r)))
        return r.value

    def is_treelet_root(self, addr):
        rc = self.L.vsrt_is_treelet_root(self.h, addr)
        if rc < 0:
            self._ck(rc)
        return bool(rc)

    def metadata_idx(self, root):
        i = c_u32()
        self._ck(self.L.vsrt_treelet_metadata_idx(self.h, root, ctypes.byref(i)))
        return i.value

    # ---- traversal ----------------------------------------------------------------------------------------
    def trace(self, mode, rays, want_trace=True, capacity=None):
        """Host-buffer call (what a reference-side caller makes): rays in, hits + CSR trace + treelet ids out."""
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY)
        n = len(rays)
        hits = np.zeros(n, _abi.HIT); offs = np.zeros(n + 1, np.uint64)
        total = c_u64()
        if not want_trace:
            self._ck(self.L.vsrt_trace_rays(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(offs), None, 0, None, ctypes.byref(total)))
            return {"hits": hits, "offsets": offs, "total": total.value}
        cap = capacity if capacity is not None else 0
        txns = np.zeros(cap, _abi.TXN); tids = np.zeros(cap, np.uint64)
        rc = self._ck(self.L.vsrt_trace_rays(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(offs),
                                             _abi.ptr(txns) if cap else None, cap, _abi.ptr(tids) if cap else None, ctypes.byref(total)),
                      allow=(-4,))
        if total.value > cap or rc == -4 or cap == 0:
            txns = np.zeros(total.value, _abi.TXN); tids = np.zeros(total.value, np.uint64)
            if total.value:
                self._ck(self.L.vsrt_trace_fetch(self.h, _abi.ptr(txns), total.value, _abi.ptr(tids)))
        return {"hits": hits, "offsets": offs, "txns": txns[:total.value], "treelet_ids": tids[:total.value]}

    def trace_into(self, mode, n, rays_ptr, hits_ptr, offsets_ptr, txns_ptr, txn_capacity, tids_ptr):
        """vsrt_trace_rays on caller-owned (e.g. pinned) host buffers given as raw addresses; returns #records."""
        total = c_u64()
        self._ck(self.L.vsrt_trace_rays(self.h, self.tlas, mode, n, rays_ptr, hits_ptr, offsets_ptr, txns_ptr, txn_capacity, tids_ptr,
                                        ctypes.byref(total)))
        return total.value

    def trace_warp(self, rays32, active_mask=0xffffffff, capacity=32 * 1024):
        rays32 = np.ascontiguousarray(rays32, dtype=_abi.RAY)
        assert len(rays32) == 32
        hits = np.zeros(32, _abi.HIT); counts = np.zeros(32, np.uint32); txns = np.zeros(capacity, _abi.TXN); total = c_u64()
        self._ck(self.L.vsrt_trace_ray_warp(self.h, self.tlas, active_mask, _abi.ptr(rays32), _abi.ptr(hits), _abi.ptr(counts), _abi.ptr(txns),
                                            capacity, ctypes.byref(total)))
        return {"hits": hits, "counts": counts, "txns": txns[:total.value]}

    def trace_device(self, mode, rays_dev_ptr, n, stream=None):
        total = c_u64()
        self._ck(self.L.vsrt_trace_rays_device(self.h, self.tlas, mode, n, rays_dev_ptr, stream, ctypes.byref(total)))
        return total.value

    def device_results(self):
        r = _abi.DeviceResults()
        self._ck(self.L.vsrt_trace_device_results(self.h, ctypes.byref(r)))
        return r

    # ---- RT-unit replay helpers ---------------------------------------------------------------------------
    def fetch_trace(self):
        """Records and 64-bit treelet ids of the last batch as they are on the device now (sorted if sort_trace ran)."""
        n = self.device_results().n_txn
        txns = np.zeros(n, _abi.TXN); tids = np.zeros(n, np.uint64)
        if n:
            self._ck(self.L.vsrt_trace_fetch(self.h, _abi.ptr(txns), n, _abi.ptr(tids)))
        return txns, tids

    def fetch_packed(self):
        """Packed records (slot << 3 | code) and treelet indices of the last batch, in traversal order."""
        n = self.device_results().n_txn
        rec = np.zeros(n, np.uint32); tix = np.zeros(n, np.uint32)
        if n:
            self._ck(self.L.vsrt_trace_fetch_packed(self.h, _abi.ptr(rec), n, _abi.ptr(tix)))
        return rec, tix

    def packed_layout(self):
        lay = _abi.PackedLayout()
        self._ck(self.L.vsrt_packed_layout_get(self.h, self.tlas, ctypes.byref(lay)))
        return lay

    def unpack(self, records, layout=None):
        """vsrt_unpack_txn over an array of packed records."""
        lay = layout or self.packed_layout()
        rec = np.ascontiguousarray(records, np.uint32); out = np.zeros(len(rec), _abi.TXN)
        self.L.vsrt_unpack_txns(ctypes.byref(lay), _abi.ptr(rec), len(rec), _abi.ptr(out))
        return out

    def node_treelet_table(self):
        """Treelet index of every slot of the packed arena (0xFFFFFFFF = none): treelet of a packed record = table[record >> 3]."""
        n = c_u64()
        self._ck(self.L.vsrt_node_treelet_table(self.h, None, 0, ctypes.byref(n)))
        t = np.zeros(n.value, np.uint32)
        self._ck(self.L.vsrt_node_treelet_table(self.h, _abi.ptr(t), n.value, ctypes.byref(n)))
        return t

    def trace_packed(self, mode, rays, want_index=False):
        """vsrt_trace_rays_packed on numpy buffers: (hits, offsets, packed records[, treelet indices])."""
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY)
        n = len(rays)
        hits = np.zeros(n, _abi.HIT); offs = np.zeros(n + 1, np.uint64); total = c_u64()
        self._ck(self.L.vsrt_trace_rays_packed(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(offs), None, 0, None, ctypes.byref(total)), allow=(-4,))
        rec = np.zeros(total.value, np.uint32); tix = np.zeros(total.value, np.uint32) if want_index else None
        self._ck(self.L.vsrt_trace_rays_packed(self.h, self.tlas, mode, n, _abi.ptr(rays), _abi.ptr(hits), _abi.ptr(offs), _abi.ptr(rec), total.value, _abi.ptr(tix), ctypes.byref(total)))
        return hits, offs, rec, tix

    def trace_packed_into(self, mode, n, rays_ptr, hits_ptr, offsets_ptr, rec_ptr, capacity, tix_ptr):
        """vsrt_trace_rays_packed on caller-owned (e.g. pinned) host buffers given as raw addresses; returns #records."""
        total = c_u64()
        self._ck(self.L.vsrt_trace_rays_packed(self.h, self.tlas, mode, n, rays_ptr, hits_ptr, offsets_ptr, rec_ptr, capacity, tix_ptr, ctypes.byref(total)))
        return total.value

    def sort_trace(self, method):
        """rt_unit::sort_mem_accesses over every ray of the last batch; returns the sorted (records, treelet ids)."""
        self._ck(self.L.vsrt_sort_trace(self.h, method))
        return self.fetch_trace()

    def prefetch_vote(self, group_offsets, heuristic=0, threshold=0.0, ray_ids=None, front=None, load_metadata=False, metadata_base=0):
        cfg = _abi.PrefetchConfig(heuristic, 1 if load_metadata else 0, threshold, metadata_base)
        go = np.ascontiguousarray(group_offsets, np.uint64)
        ids = None if ray_ids is None else np.ascontiguousarray(ray_ids, np.uint64)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        dec = np.zeros(len(go) - 1, _abi.PDEC)
        self._ck(self.L.vsrt_prefetch_vote(self.h, ctypes.byref(cfg), len(go) - 1, _abi.ptr(go), _abi.ptr(ids), _abi.ptr(fr), _abi.ptr(dec)))
        return dec

    def prefetch_chunks(self, decisions, heuristic=0, load_metadata=False, metadata_base=0):
        cfg = _abi.PrefetchConfig(heuristic, 1 if load_metadata else 0, 0.0, metadata_base)
        dec = np.ascontiguousarray(decisions, _abi.PDEC)
        offs = np.zeros(len(dec) + 1, np.uint64); n = c_u64()
        self._ck(self.L.vsrt_prefetch_chunks(self.h, ctypes.byref(cfg), len(dec), _abi.ptr(dec), _abi.ptr(offs), None, None, 0, ctypes.byref(n)), allow=(-4,))
        ca = np.zeros(n.value, np.uint64); co = np.zeros(n.value, np.uint64)
        if n.value:
            self._ck(self.L.vsrt_prefetch_chunks(self.h, ctypes.byref(cfg), len(dec), _abi.ptr(dec), _abi.ptr(offs), _abi.ptr(ca), _abi.ptr(co), n.value, ctypes.byref(n)))
        return offs, ca, co

    def table_events(self, tid_x=None, want_anyhit=True):
        """Baseline intersection / any-hit table calls of the last batch: (CSR offsets, events, any-hit Hit_data)."""
        n = self.device_results().n_rays
        tx = None if tid_x is None else np.ascontiguousarray(tid_x, np.uint8)
        offs = np.zeros(n + 1, np.uint64); tot = c_u64()
        self._ck(self.L.vsrt_table_events(self.h, _abi.ptr(tx), _abi.ptr(offs), None, None, 0, ctypes.byref(tot)), allow=(-4,))
        ev = np.zeros(tot.value, _abi.TEV); ah = np.zeros(tot.value, _abi.HIT) if want_anyhit else None
        if tot.value:
            self._ck(self.L.vsrt_table_events(self.h, _abi.ptr(tx), _abi.ptr(offs), _abi.ptr(ev), _abi.ptr(ah), tot.value, ctypes.byref(tot)))
        return offs, ev, ah

    def table_event_stores(self, events, table_bases):
        """The two MemoryStoreTransactionRecords of every event; table_bases = (intersection_table, anyhit_table) addresses."""
        out = np.zeros((len(events), 2), _abi.STORE)
        for i in range(len(events)):
            self.L.vsrt_table_event_stores(events[i:i + 1].ctypes.data_as(c_vp), int(table_bases[int(events[i]["table"])]), out[i].ctypes.data_as(c_vp))
        return out

    def coalescing_events(self, event_offsets, events):
        """Function_Call_Coalescing intersection table replayed over the table-0 events of table_events(): one
        vsrt_coalescing_event (row, appended, n_loads, first_new_load) per event."""
        offs = np.ascontiguousarray(event_offsets, np.uint64); ev = np.ascontiguousarray(events, _abi.TEV)
        out = np.zeros(len(ev), _abi.CEV)
        self._ck(self.L.vsrt_coalescing_events(self.h, len(offs) - 1, _abi.ptr(offs), _abi.ptr(ev), _abi.ptr(out)))
        return out

    def coalescing_trace(self, offsets, txns, event_offsets, events, cev, table_base):
        """The per-ray transaction and store lists with a Coalescing intersection table at table_base: the load records of
        every call spliced in after its PROCEDURAL_LEAF record (rows first_new_load .. n_loads-1), stores in call order.
        Returns (txn_offsets, txns, store_offsets, stores)."""
        offsets = np.asarray(offsets, np.int64); eo = np.asarray(event_offsets, np.int64)
        out_t, out_s, to, so = [np.zeros(0, _abi.TXN)], [np.zeros(0, _abi.STORE)], [0], [0]
        one = np.zeros(1, _abi.TXN); st3 = np.zeros(3, _abi.STORE)
        for r in range(len(offsets) - 1):
            seg = txns[offsets[r]:offsets[r + 1]]
            pos, nt, ns = 0, 0, 0
            for k in range(eo[r], eo[r + 1]):
                if events[k]["table"] != 0:
                    continue
                rec = int(events[k]["record"])
                out_t.append(seg[pos:rec + 1]); nt += rec + 1 - pos; pos = rec + 1
                for row in range(int(cev[k]["first_new_load"]), int(cev[k]["n_loads"])):
                    self.L.vsrt_coalescing_event_load(row, int(table_base), one.ctypes.data_as(c_vp)); out_t.append(one.copy()); nt += 1
                n = self.L.vsrt_coalescing_event_stores(events[k:k + 1].ctypes.data_as(c_vp), cev[k:k + 1].ctypes.data_as(c_vp), int(table_base), st3.ctypes.data_as(c_vp))
                out_s.append(st3[:n].copy()); ns += n
            out_t.append(seg[pos:]); nt += len(seg) - pos
            to.append(to[-1] + nt); so.append(so[-1] + ns)
        return np.array(to, np.uint64), np.concatenate(out_t), np.array(so, np.uint64), np.concatenate(out_s)

    def schedule_pick(self, scheduler, unit_warp_offsets, warp_ray_ids, stalled=None, last_prefetched=None, front=None):
        uo = np.ascontiguousarray(unit_warp_offsets, np.uint64); ids = np.ascontiguousarray(warp_ray_ids, np.uint64)
        st = None if stalled is None else np.ascontiguousarray(stalled, np.uint8)
        lp = None if last_prefetched is None else np.ascontiguousarray(last_prefetched, np.uint64)
        fr = None if front is None else np.ascontiguousarray(front, np.uint32)
        pick = np.zeros(len(uo) - 1, np.int64)
        self._ck(self.L.vsrt_schedule_pick(self.h, scheduler, len(uo) - 1, _abi.ptr(uo), _abi.ptr(ids), _abi.ptr(st), _abi.ptr(lp), _abi.ptr(fr), _abi.ptr(pick)))
        return pick

    # ---- counters -----------------------------------------------------------------------------------------
    def counters(self):
        a = np.zeros(_abi.N_SUM + _abi.N_MAX, np.uint64)
        self._ck(self.L.vsrt_get_counters(self.h, _abi.ptr(a)))
        return dict(zip(_abi.COUNTER_FIELDS, (int(x) for x in a)))

    def reset_counters(self):
        self._ck(self.L.vsrt_reset_counters(self.h))

    def treelet_histogram(self):
        n = self.treelet_info().n_treelets
        h = np.zeros(n, np.uint64)
        self._ck(self.L.vsrt_get_treelet_histogram(self.h, _abi.ptr(h), n))
        return h

    # ---- multi-GPU counter / histogram reduce (NCCL inside the library) ----------------------------------
    def comm_init(self, n_ranks, rank, unique_id):
        b = (ctypes.c_uint8 * _abi.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.vsrt_comm_init(self.h, n_ranks, rank, b))

    def comm_destroy(self):
        self._ck(self.L.vsrt_comm_destroy(self.h))

    def reduce_counters(self, stream=None):
        """Enqueue one reduce of what this rank traced since the previous one (returns without waiting)."""
        self._ck(self.L.vsrt_reduce_counters(self.h, stream))

    def reduce_wait(self, stream=None, host=False):
        self._ck(self.L.vsrt_reduce_wait(self.h, c_vp(-1 & 0xFFFFFFFFFFFFFFFF) if host else stream))

    def reduced(self, want_hist=True):
        """Global totals over all ranks: (counter dict, treelet histogram)."""
        a = np.zeros(_abi.N_SUM + _abi.N_MAX, np.uint64)
        n = self.treelet_info().n_treelets if want_hist else 0
        h = np.zeros(n, np.uint64) if want_hist else None
        self._ck(self.L.vsrt_reduced_get(self.h, _abi.ptr(a), _abi.ptr(h), n))
        return dict(zip(_abi.COUNTER_FIELDS, (int(x) for x in a))), h

    def tb_stats(self):
        """Statistics of the last batch traced by the treelet-binned kernel (VSRT_K1_TB=1); debug export, not part of vsrt.h."""
        a = (ctypes.c_ulonglong * 8)()
        self.L.vsrt_debug_tb_stats.argtypes = [c_vp, c_vp]
        self._ck(self.L.vsrt_debug_tb_stats(self.h, a))
        names = ("rays_processed", "rays_in_staged_treelet", "ctas_staged", "bytes_staged", "visits_from_smem", "visits_from_arena", "rounds")
        return {n: int(a[i]) for i, n in enumerate(names)}

    def enable_node_histogram(self, on=True):
        self._ck(self.L.vsrt_enable_node_histogram(self.h, 1 if on else 0))

    def node_histogram(self, reduced=False):
        """Records per node address, per slot of the packed arena (this rank's, or the global one after reduce_counters)."""
        n = c_u64()
        self._ck(self.L.vsrt_get_node_histogram(self.h, None, 0, ctypes.byref(n)))
        h = np.zeros(n.value, np.uint64)
        if reduced:
            self._ck(self.L.vsrt_reduced_get_node_histogram(self.h, _abi.ptr(h), n.value))
        else:
            self._ck(self.L.vsrt_get_node_histogram(self.h, _abi.ptr(h), n.value, ctypes.byref(n)))
        return h

    def counters_device(self):
        cp, hp, n = c_vp(), c_vp(), c_u64()
        self._ck(self.L.vsrt_counters_device(self.h, ctypes.byref(cp), ctypes.byref(hp), ctypes.byref(n)))
        return cp.value, hp.value, n.value

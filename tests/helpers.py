"""Shared builders for the tests: hand-made arenas (struct.pack), seeded ray sets, comparison helpers."""
import struct
import numpy as np
from vsrt import scene as sc, _abi


def mk_header(root_off, mn, mx):
    b = bytearray(64)
    struct.pack_into("<Q6f", b, 0, root_off, *mn, *mx)
    return bytes(b)


def mk_internal(origin, child_off, exps, infos, lo, hi):
    b = bytearray(64)
    struct.pack_into("<3fi", b, 0, *origin, child_off)
    b[18:21] = bytes([e & 0xff for e in exps])
    b[21] = 0xff
    for i, inf in enumerate(infos):
        b[22 + i] = inf
    for i, (l, h) in enumerate(zip(lo, hi)):
        b[28 + i] = l[0]; b[34 + i] = h[0]; b[40 + i] = l[1]; b[46 + i] = h[1]; b[52 + i] = l[2]; b[58 + i] = h[2]
    return bytes(b)


def mk_instance(bvh_off, inst_id):
    b = bytearray(128)
    ident = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
    struct.pack_into("<12f", b, 16, *ident)
    struct.pack_into("<QII", b, 64, bvh_off & 0xFFFFFFFFFFFFFFFF, inst_id, 0)
    struct.pack_into("<12f", b, 80, *ident)
    return bytes(b)


def mk_quad(prim, geom, v, delta=0):
    b = bytearray(64)
    struct.pack_into("<II", b, 0, 0xff000000, geom)
    struct.pack_into("<II", b, 8, prim, delta)
    struct.pack_into("<9f", b, 16, *[c for p in v for c in p])
    return bytes(b)


def kat_arena(z448=1.0, z512=2.0):
    """The worked vector of SURVEY.md section 8(c): TLAS hdr@0, TLAS internal@64, instance leaf@192,
    BLAS hdr@320, BLAS internal@384, quads @448 (prim 100) and @512 (prim 101)."""
    a = mk_header(64, (-8,) * 3, (8,) * 3)
    a += mk_internal((-8, -8, -8), 2, (12, 12, 12), [6, 0, 0, 0, 0, 0], [(0, 0, 0)] * 6, [(1, 1, 1)] + [(0, 0, 0)] * 5)
    a += b"\0" * 64
    a += mk_instance(128, 7)
    a += mk_header(64, (-8,) * 3, (8,) * 3)
    a += mk_internal((-8, -8, -8), 1, (12, 12, 12), [17, 17, 0, 0, 0, 0], [(0, 0, 0)] * 6, [(1, 1, 1), (1, 1, 1)] + [(0, 0, 0)] * 4)
    a += mk_quad(100, 3, [(-1, -1, z448), (1, -1, z448), (0, 1, z448)])
    a += mk_quad(101, 3, [(-1, -1, z512), (1, -1, z512), (0, 1, z512)])
    return sc.Arena(a, 0, [(320, len(a) - 320)])


def kat_arena_unordered(z448=1.0, z512=2.0):
    """kat_arena with the quantised x and z bounds of the BLAS node's second child swapped (lower 1, upper 0): a box the
    reference still intersects (its slab test orders the planes with min / max, vulkan_ray_tracing.cc:183-257) but whose
    near / far planes cannot be read off the ray's direction signs."""
    a = kat_arena(z448, z512)
    for ax in (0, 2):
        a.bytes[384 + 28 + 12 * ax + 1] = 1      # lower bound of child 1 on this axis
        a.bytes[384 + 34 + 12 * ax + 1] = 0      # upper bound
    return a


def kat_ray(flags):
    r = np.zeros(1, _abi.RAY)
    r["origin"][0] = (0, 0, -5); r["direction"][0] = (0, 0, 1); r["tmin"] = 0; r["tmax"] = 100; r["ray_flags"] = flags
    return r


def mixed_rays(n_random, seed, w=48, h=32):
    """Primary + incoherent rays with a mix of flags, a raised tmin and exactly-zero direction components."""
    r = np.concatenate([sc.rays_primary(w, h, flags=0), sc.rays_random(n_random, seed=seed)])
    i = np.arange(len(r))
    r["ray_flags"] = np.where(i % 3 == 0, 1, 0) | np.where(i % 5 == 0, 4, 0)
    r["tmin"][::7] = 0.5
    r["direction"][::11, 0] = 0.0
    r["tmax"][::13] = 2.0
    return r


def assert_trace_equal(o, g, what="", skip_unknown_tids=False):
    """o: oracle result (OHIT hits), g: CUDA result (vsrt_hit hits).  Bit-exact on everything."""
    assert np.array_equal(o["offsets"], g["offsets"]), what + ": per-ray record counts differ"
    assert np.array_equal(o["txns"]["address"], g["txns"]["address"]), what + ": node-visit addresses differ"
    assert np.array_equal(o["txns"]["size"], g["txns"]["size"]) and np.array_equal(o["txns"]["type"], g["txns"]["type"]), what + ": record size/type differ"
    if skip_unknown_tids:
        # traceRay with per-BLAS device offsets emits BLAS-node addresses the reference's address-keyed treelet map
        # does not contain (its addrToTreeletID would assert, :470); the oracle reports ~0 there
        known = o["treelet_ids"] != np.uint64(0xFFFFFFFFFFFFFFFF)
        assert np.array_equal(o["treelet_ids"][known], g["treelet_ids"][known]), what + ": treelet ids differ"
    else:
        assert np.array_equal(o["treelet_ids"], g["treelet_ids"]), what + ": treelet ids differ"
    oh, gh = o["hits"], g["hits"]
    assert np.array_equal(oh["hit"], gh["hit_geometry"]), what + ": hit flags differ"
    assert np.array_equal(oh["prim"], gh["primitive_index"]), what + ": primitive ids differ"
    assert np.array_equal(oh["geom"], gh["geometry_index"]) and np.array_equal(oh["instance_id"], gh["instance_index"]), what + ": geometry/instance ids differ"
    assert np.array_equal(oh["n_all_hits"], gh["n_all_hits"]), what + ": n_all_hits differ"
    # north_star asks for 1e-6 relative on t and barycentrics; the implementation is bit-exact, so test that
    assert np.array_equal(oh["t"].view(np.uint32), gh["world_min_thit"].view(np.uint32)), what + ": hit t differs bitwise"
    assert np.array_equal(oh["bary"].view(np.uint32), gh["barycentric"].view(np.uint32)), what + ": barycentrics differ bitwise"
    assert np.array_equal(oh["point"].view(np.uint32), gh["intersection_point"].view(np.uint32)), what + ": intersection points differ bitwise"


def assert_tables_equal(to, tg, what=""):
    for k in ("roots", "counts", "meta_idx", "node_addr", "node_size", "map_nodes", "map_roots"):
        assert np.array_equal(to[k], tg[k]), "%s: treelet table field %s differs" % (what, k)


COUNTER_MAP = {"num_hits": "num_hits", "num_any_hits": "num_any_hits", "n_anyhit_rays": "n_anyhit_rays",
               "n_closesthit_rays": "n_closesthit_rays", "max_nodes_per_ray": "max_nodes_per_ray",
               "tot_nodes_per_ray": "tot_nodes_per_ray", "max_tree_depth": "max_tree_depth", "ray_count": "ray_count"}


COALESCING_ENTRY = 292   # sizeof(Coalescing_Entry): u32 hitGroupIndex, bool thread_mask[32], {u32, u32} shader_data[32]


def coalescing_splice(offsets, txns, ev_offsets, ev_table, ev_tid, cev, base=1 << 63):
    """Per-ray transaction / store lists with a Coalescing intersection table at `base`, built from a plain trace, the table
    events of its rays (table-0 events are the rays' PROCEDURAL_LEAF records, in order) and the replayed table (cev):
    after the j-th PROCEDURAL_LEAF record come the load records of rows first_new_load .. n_loads-1
    (vulkan_ray_tracing.cc:2186-2200); stores per call: [hitGroupIndex if appended], thread_mask[tid], shader_data[tid]
    (intersection_table.cc:73-74, :88-90)."""
    offsets = np.asarray(offsets, np.int64); eo = np.asarray(ev_offsets, np.int64)
    t_out, s_out, to, so = [], [], [0], [0]
    for r in range(len(offsets) - 1):
        seg = txns[offsets[r]:offsets[r + 1]]
        ks = [k for k in range(eo[r], eo[r + 1]) if ev_table[k] == 0]
        procs = np.flatnonzero(seg["type"] == 6)
        assert len(procs) == len(ks)
        pos = 0
        for j, k in enumerate(ks):
            for rec in seg[pos:procs[j] + 1]:
                t_out.append((int(rec["address"]), int(rec["size"]), int(rec["type"])))
            pos = procs[j] + 1
            for row in range(int(cev[k]["first_new_load"]), int(cev[k]["n_loads"])):
                t_out.append((base + row * COALESCING_ENTRY, 4, 7))
            rb = base + int(cev[k]["row"]) * COALESCING_ENTRY
            if cev[k]["appended"]:
                s_out.append((rb, 4, 0))
            s_out.append((rb + 4 + int(ev_tid[k]), 1, 0)); s_out.append((rb + 36 + 8 * int(ev_tid[k]), 8, 0))
        for rec in seg[pos:]:
            t_out.append((int(rec["address"]), int(rec["size"]), int(rec["type"])))
        to.append(len(t_out)); so.append(len(s_out))
    return (np.array(to, np.uint64), np.array(t_out, dtype=_abi.TXN) if t_out else np.zeros(0, _abi.TXN),
            np.array(so, np.uint64), np.array(s_out, dtype=_abi.STORE) if s_out else np.zeros(0, _abi.STORE))


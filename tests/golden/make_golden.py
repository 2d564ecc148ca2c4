"""Generates the golden fixtures in this directory from the REFERENCE'S OWN CODE (oracle/_ref/libvsrt_ref.so,
built by oracle/build_ref.sh from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the exact arena bytes, the rays and what the reference returned (treelet tables, traces,
treelet ids, hit records, counters) with addresses stored RELATIVE to the arena base, so they can be replayed at
any host address.  The reference ships no tests or vectors for this path; these files are
the pin for the oracle restatement and for the CUDA path on the GPU box, where /root/reference does not exist."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import __graft_entry__ as g   # noqa: E402
g.build_cpu()
from vsrt import scene as sc  # noqa: E402
import helpers                # noqa: E402
import oracles                # noqa: E402

CASES = [
    # name, arena factory, rays, budgets, delta
    ("kat", lambda: helpers.kat_arena(2.0, 1.0), lambda: np.concatenate([helpers.kat_ray(1), helpers.kat_ray(0), helpers.kat_ray(4)]), (256, 512), 0),
    ("soup300", lambda: sc.Scene(300, seed=300), lambda: helpers.mixed_rays(300, 1, 16, 12), (192, 512, 49152), 0),
    ("inst2k", lambda: sc.Scene(2000, seed=2000, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS | sc.F_HOLES), lambda: helpers.mixed_rays(400, 2, 16, 12), (256, 4096), 0),
    ("inst2k_off", lambda: sc.Scene(2000, seed=2001, n_blas=2, n_instances=4, flags=sc.F_TRANSFORMS), lambda: helpers.mixed_rays(300, 3, 16, 12), (512,), 0x100000),
]


def rel(a, base):
    return (a.astype(np.uint64) - np.uint64(base)).astype(np.uint64)


def main():
    ref = oracles.RefOracle()
    for name, mk_arena, mk_rays, budgets, delta in CASES:
        arena = mk_arena(); rays = mk_rays()
        out = {"arena": np.array(arena.bytes), "tlas_offset": np.uint64(arena.tlas_offset), "blas": np.array(arena.blas, np.uint64).reshape(-1, 2),
               "rays": rays, "budgets": np.array(budgets, np.uint32), "delta": np.uint64(delta)}
        for b in budgets:
            ref.register(arena, delta); ref.form(b)
            t = ref.tables()
            for k in ("roots", "node_addr", "map_nodes", "map_roots"):
                out["b%d_%s" % (b, k)] = rel(t[k], arena.base)
            for k in ("counts", "meta_idx", "node_size"):
                out["b%d_%s" % (b, k)] = t[k]
            for mode in (0, 1):
                r = ref.trace(mode, rays)
                p = "b%d_m%d_" % (b, mode)
                out[p + "offsets"] = r["offsets"]; out[p + "addr"] = rel(r["txns"]["address"], arena.base)
                out[p + "size"] = r["txns"]["size"].astype(np.uint8); out[p + "type"] = r["txns"]["type"].astype(np.uint8)
                out[p + "tid"] = rel(r["treelet_ids"], arena.base); out[p + "hits"] = r["hits"]
            c = ref.counters()
            out["b%d_counters" % b] = np.array([c[k] for k in oracles.OCNT_FIELDS], np.uint64)
        path = os.path.join(HERE, "golden_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%s: %d bytes (arena %d B, %d rays)" % (path, os.path.getsize(path), arena.size, len(rays)))


def replay_groups(n_rays):
    """The groups / units every replay fixture test uses (deterministic)."""
    rng = np.random.default_rng(77)
    go = np.array([0, 32, 64, 200, n_rays], np.uint64)
    front = rng.integers(0, 9, n_rays).astype(np.uint32)
    unit_offs = np.array([0, 3, 4, 9], np.uint64)
    lanes = rng.integers(0, n_rays, 9 * 32).astype(np.uint64)
    lanes[rng.random(len(lanes)) < 0.15] = np.uint64(0xFFFFFFFFFFFFFFFF)
    stalled = np.array([0, 1, 0, 0, 1, 0, 0, 0, 1], np.uint8)
    return go, front, unit_offs, lanes, stalled


def main_replay():
    """golden_replay.npz: what the reference's rt_unit bodies (sort_mem_accesses, the prefetch vote block, schedule_next_warp)
    and its -remap_to_treelet_layout traversal return on one instanced scene (shared BLAS)."""
    ref = oracles.RefOracle()
    arena = sc.Scene(1500, seed=1500, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    rays = helpers.mixed_rays(320, 4, 16, 12)
    budget, stride = 1024, 128
    out = {"arena": np.array(arena.bytes), "tlas_offset": np.uint64(arena.tlas_offset), "blas": np.array(arena.blas, np.uint64).reshape(-1, 2),
           "rays": rays, "budget": np.uint32(budget), "stride": np.uint32(stride)}
    ref.register(arena); ref.form(budget)
    go, front, unit_offs, lanes, stalled = replay_groups(len(rays))
    for mode in (0, 1):
        t = ref.trace(mode, rays)
        for method in (0, 1):
            st = ref.sort_trace(method, t)
            p = "m%d_s%d_" % (mode, method)
            out[p + "addr"] = rel(st["address"], arena.base); out[p + "size"] = st["size"].astype(np.uint8); out[p + "type"] = st["type"].astype(np.uint8)
    t = ref.trace(1, rays)
    meta = (0x5000000000, (budget // 64) * 4)
    for h, thr in ((0, 0.0), (1, 0.4), (2, 0.0), (3, 0.0)):
        for use_meta in (0, 1):
            decs = []; chunks = []; owners = []; offs = [0]
            for g in range(len(go) - 1):
                d, ca, co = ref.prefetch_vote(t, np.arange(go[g], go[g + 1]), h, thr, front, meta if use_meta else None)
                d = d.copy()
                if d["root"]:
                    d["root"] = d["root"] - np.uint64(arena.base) + np.uint64(1)       # 0 stays "no root"; the first treelet sits at offset 0
                decs.append(d)
                node = co < np.uint64(meta[0])          # metadata rows live at meta[0]; node chunks are arena addresses
                chunks.append(np.where(node, ca - np.uint64(arena.base), ca)); owners.append(np.where(node, co - np.uint64(arena.base), co))
                offs.append(offs[-1] + len(ca))
            p = "h%d_meta%d_" % (h, use_meta)
            out[p + "dec"] = np.array(decs, oracles.PDEC); out[p + "chunk_off"] = np.array(offs, np.uint64)
            out[p + "chunk_addr"] = np.concatenate(chunks).astype(np.uint64); out[p + "chunk_owner"] = np.concatenate(owners).astype(np.uint64)
    roots = np.unique(t["treelet_ids"])
    lp = np.array([roots[0], 0, roots[min(2, len(roots) - 1)]], np.uint64)
    out["sched_lp"] = np.where(lp != 0, lp - np.uint64(arena.base) + np.uint64(1), 0).astype(np.uint64)   # offset + 1, 0 = none
    for sched in (0, 1, 2):
        picks = []
        for u in range(len(unit_offs) - 1):
            w0, w1 = int(unit_offs[u]), int(unit_offs[u + 1])
            picks.append(ref.schedule_pick(t, sched, int(lp[u]), lanes[32 * w0:32 * w1], stalled[w0:w1], front))
        out["sched%d_pick" % sched] = np.array(picks, np.int64)
    # -remap_to_treelet_layout 1 (addresses relative to treelet_layout_bvh).  Its own scene, one instance per BLAS: on a
    # shared BLAS the reference aborts in remapBVHToTreeletLayout (assert at vulkan_ray_tracing.cc:1487).
    arena2 = sc.Scene(1500, seed=1501, n_blas=2, n_instances=2, flags=sc.F_TRANSFORMS)
    out["remap_arena"] = np.array(arena2.bytes); out["remap_tlas_offset"] = np.uint64(arena2.tlas_offset)
    out["remap_blas"] = np.array(arena2.blas, np.uint64).reshape(-1, 2)
    ref.register(arena2, remap=True, stride=stride); ref.form(budget)
    base = ref.remap_table()[0]
    for mode in (0, 1):
        r = ref.trace(mode, rays)
        p = "remap_m%d_" % mode
        out[p + "offsets"] = r["offsets"]; out[p + "addr"] = rel(r["txns"]["address"], base)
        out[p + "size"] = r["txns"]["size"].astype(np.uint8); out[p + "type"] = r["txns"]["type"].astype(np.uint8); out[p + "tid"] = rel(r["treelet_ids"], base)
    path = os.path.join(HERE, "replay_inst1500.npz")
    np.savez_compressed(path, **out)
    print("%s: %d bytes" % (path, os.path.getsize(path)))


def main_tables():
    """tables_proc1200.npz: the reference's Baseline intersection / any-hit tables (intersection_table.cc) in the loop of
    both traversals on a scene with procedural leaves; rays [32g, 32g+32) = one CTA, tid = r % 32."""
    ref = oracles.RefOracle()
    arena = sc.Scene(1200, seed=1200, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(600, 6, 16, 12)
    out = {"arena": np.array(arena.bytes), "tlas_offset": np.uint64(arena.tlas_offset), "blas": np.array(arena.blas, np.uint64).reshape(-1, 2),
           "rays": rays, "budget": np.uint32(512)}
    ref.register(arena); ref.form(512)
    ib, ab = ref.table_bases()
    for mode in (0, 1):
        counts, ev, ah = ref.table_events(mode, rays)
        p = "m%d_" % mode
        out[p + "counts"] = counts
        for k in ("table", "shader_counter", "hit_group_index", "primitive_id", "instance_id", "tid", "store_size"):
            out[p + k] = ev[k]
        base = np.where(ev["table"] == 1, np.uint64(ab), np.uint64(ib))
        out[p + "store_off"] = ev["store_addr"] - base[:, None]             # offsets inside the table
        out[p + "anyhit"] = ah
    path = os.path.join(HERE, "tables_proc1200.npz")
    np.savez_compressed(path, **out)
    print("%s: %d bytes" % (path, os.path.getsize(path)))


def main_coalescing():
    """coalescing_proc1500.npz: the reference's traceRay with its own Function_Call_Coalescing intersection table in the loop
    (intersection_table.cc:43-98; a fresh table per 32 rays, tid = r % 32): per-ray transaction lists with the
    Intersection_Table_Load records merged in, and store lists.  Table addresses are offsets from the table base | 1 << 63,
    BVH addresses are relative to the arena base."""
    ref = oracles.RefOracle()
    arena = sc.Scene(1500, seed=8, n_blas=2, n_instances=5, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(800, 12, 24, 16)
    out = {"arena": np.array(arena.bytes), "tlas_offset": np.uint64(arena.tlas_offset), "blas": np.array(arena.blas, np.uint64).reshape(-1, 2),
           "rays": rays, "budget": np.uint32(512)}
    ref.register(arena); ref.form(512)
    to, tx, so, st, entry = ref.trace_coalescing(0, rays)
    tab = (tx["address"] >> np.uint64(63)) != 0
    out["txn_offsets"] = to; out["txn_addr"] = np.where(tab, tx["address"], tx["address"] - np.uint64(arena.base))
    out["txn_size"] = tx["size"].astype(np.uint8); out["txn_type"] = tx["type"].astype(np.uint8)
    out["store_offsets"] = so; out["store_addr"] = st["address"]; out["store_size"] = st["size"].astype(np.uint8)
    out["entry_size"] = np.uint32(entry)
    assert int(tab.sum()) > 25 and (st["address"] >> np.uint64(63)).all()
    path = os.path.join(HERE, "coalescing_proc1500.npz")
    np.savez_compressed(path, **out)
    print("%s: %d bytes (%d table loads, %d stores)" % (path, os.path.getsize(path), int(tab.sum()), len(st)))


if __name__ == "__main__":
    main()
    main_replay()
    main_tables()
    main_coalescing()

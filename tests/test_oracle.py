"""CPU tests of the oracle itself (no GPU): the plain-C restatement against (1) the golden fixtures recorded
from the reference's own code and (2), when oracle/_ref was built here, the compiled reference live."""
import numpy as np
import pytest
from vsrt import scene as sc, _abi
import helpers
import oracles
import golden_util


def same_hits(a, b):
    return all(np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)) for f in a.dtype.names)


@pytest.mark.parametrize("path", golden_util.fixtures(), ids=lambda p: p.split("golden_")[-1][:-4])
def test_port_matches_golden(path):
    z, arena, rays = golden_util.load(path)
    delta = int(z["delta"])
    port = oracles.PortOracle()
    for b in (int(x) for x in z["budgets"]):
        port.register(arena, delta); port.form(b)
        exp = golden_util.expected_tables(z, b, arena.base)
        got = port.tables()
        for k in exp:
            assert np.array_equal(exp[k], got[k]), "budget %d table %s" % (b, k)
        for mode in (0, 1):
            e = golden_util.expected_trace(z, b, mode, arena.base)
            g = port.trace(mode, rays)
            assert np.array_equal(e["offsets"], g["offsets"])
            assert np.array_equal(e["txns"], g["txns"])
            assert np.array_equal(e["treelet_ids"], g["treelet_ids"])
            assert same_hits(e["hits"], g["hits"])
        c = port.counters()
        assert [c[k] for k in oracles.OCNT_FIELDS] == [int(x) for x in z["b%d_counters" % b]]


def test_golden_fixtures_exist():
    assert len(golden_util.fixtures()) >= 4


def test_kat_vector_port():
    """SURVEY.md 8(c) worked vector on the port oracle."""
    a = helpers.kat_arena(2.0, 1.0)
    port = oracles.PortOracle(); port.register(a); port.form(256)
    t = port.tables()
    assert [int(r) - a.base for r in t["roots"]] == [0, 192, 448, 512] and list(t["counts"]) == [2, 3, 1, 1]
    d = port.trace(0, helpers.kat_ray(1)); tr = port.trace(1, helpers.kat_ray(0))
    assert int(d["hits"]["prim"][0]) == 100 and int(tr["hits"]["prim"][0]) == 101
    assert int(d["txns"]["type"][-1]) == 5 and int(tr["txns"]["type"][-1]) == 4
    assert float(d["hits"]["t"][0]) == 6.0 and float(tr["hits"]["t"][0]) == 6.0


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("ntri,nb,ni,flags,fan,delta", [
    (300, 1, 1, 0, 6, 0),
    (2000, 2, 3, sc.F_TRANSFORMS | sc.F_HOLES, 6, 0),
    (5000, 3, 7, sc.F_TRANSFORMS, 4, 0x100000),
    (12000, 1, 2, sc.F_TRANSFORMS | sc.F_HOLES, 6, 0),
])
def test_port_matches_reference_live(ntri, nb, ni, flags, fan, delta):
    s = sc.Scene(ntri, seed=ntri + 1, n_blas=nb, n_instances=ni, flags=flags, max_fanout=fan)
    rays = helpers.mixed_rays(600, ntri)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    for budget in (192, 512, 2048, 49152):
        ref.register(s, delta); ref.form(budget); port.register(s, delta); port.form(budget)
        tr, tp = ref.tables(), port.tables()
        for k in tr:
            assert np.array_equal(tr[k], tp[k]), k
        for mode in (0, 1):
            a, b = ref.trace(mode, rays), port.trace(mode, rays)
            assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["txns"], b["txns"])
            assert np.array_equal(a["treelet_ids"], b["treelet_ids"]) and same_hits(a["hits"], b["hits"])
        assert ref.counters() == port.counters()


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_remap_matches_reference():
    s = sc.Scene(2500, seed=4, n_blas=2, n_instances=3)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s, remap=True, stride=128); ref.form(1024)
    port.register(s); port.form(1024)
    base, o, m = ref.remap_table()
    po, pm = port.remap_table(base, 128)
    assert np.array_equal(o, po) and np.array_equal(m, pm)


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_remapped_traces_match_reference():
    """-remap_to_treelet_layout 1: the port's remapped trace (visit order unchanged, addresses and treelet ids through the
    remap table) equals what the reference itself emits with the option on, in both variants."""
    s = sc.Scene(2500, seed=4, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(1500, 5)
    for mode in (0, 1):
        ref, port = oracles.RefOracle(), oracles.PortOracle()
        ref.register(s, remap=True, stride=128); ref.form(1024)
        port.register(s); port.form(1024)
        base = ref.remap_table()[0]
        a, b = ref.trace(mode, rays), port.trace_remapped(mode, rays, base, 128, 1024)
        assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["txns"], b["txns"])
        assert np.array_equal(a["treelet_ids"], b["treelet_ids"]) and same_hits(a["hits"], b["hits"])


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("budget", [512, 4096])
def test_rt_unit_sort_matches_reference(budget):
    """rt_unit::sort_mem_accesses (shader.cc:3012-3089), both -sort_method values, on a scene whose shared BLAS puts
    nodes into several treelet lists."""
    s = sc.Scene(2500, seed=4, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(400, 5)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s); ref.form(budget); port.register(s); port.form(budget)
    for mode in (0, 1):
        t = port.trace(mode, rays)
        for method in (0, 1):
            a, b = ref.sort_trace(method, t), port.sort_trace(method, t)
            assert np.array_equal(a, b), (mode, method)
            assert not np.array_equal(a, t["txns"])      # the sort does something on these traces


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_rt_unit_prefetch_vote_matches_reference():
    """The treelet-prefetch vote block of rt_unit::cycle (shader.cc:3419-3685): decision and queued 32-byte chunks for
    groups of rays at different points of their lists, all four heuristics, with and without metadata loads."""
    s = sc.Scene(2500, seed=4, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(512, 9)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s); ref.form(1024); port.register(s); port.form(1024)
    t = port.trace(1, rays)
    counts = np.diff(t["offsets"]).astype(np.int64)
    rng = np.random.default_rng(3)
    checked = 0
    for g0, g1 in ((0, 32), (32, 160), (100, 512), (7, 9)):
        ids = np.arange(g0, g1)
        for step in (0, 1, 3, 8, 10 ** 6):
            front = np.minimum(rng.integers(0, step + 1, len(rays)), 10 ** 6).astype(np.uint32)
            for h, thr in ((0, 0.0), (1, 0.3), (1, 0.9), (2, 0.0), (3, 0.0)):
                for meta in (None, (0x5000000000, (1024 // 64) * 4)):
                    (da, ca, oa), (db, cb, ob) = ref.prefetch_vote(t, ids, h, thr, front, meta), port.prefetch_vote(t, ids, h, thr, front, meta)
                    assert da == db, (g0, g1, step, h, thr, da, db)
                    assert np.array_equal(ca, cb) and np.array_equal(oa, ob)
                    checked += int(da["submit"])
    assert checked > 20


def test_port_matches_replay_fixture():
    """The restatements of the rt_unit helpers and of the remapped traversal against tests/golden/replay_inst1500.npz,
    recorded from the reference's own bodies (works without /root/reference)."""
    z, arena, arena2, rays = golden_util.load_replay()
    budget, stride = int(z["budget"]), int(z["stride"])
    port = oracles.PortOracle(); port.register(arena); port.form(budget)
    go, front, unit_offs, lanes, stalled = golden_util.replay_groups(len(rays))
    for mode in (0, 1):
        t = port.trace(mode, rays)
        for method in (0, 1):
            assert np.array_equal(port.sort_trace(method, t), golden_util.replay_txns(z, "m%d_s%d_" % (mode, method), arena.base)), (mode, method)
    t = port.trace(1, rays)
    for h, thr in ((0, 0.0), (1, 0.4), (2, 0.0), (3, 0.0)):
        for use_meta in (0, 1):
            p = "h%d_meta%d_" % (h, use_meta)
            offs, ca, co = golden_util.replay_chunks(z, p, arena.base)
            for g in range(len(go) - 1):
                d, a, b = port.prefetch_vote(t, np.arange(go[g], go[g + 1]), h, thr, front, (golden_util.META_BASE, (budget // 64) * 4) if use_meta else None)
                want = z[p + "dec"][g]
                assert int(d["root"]) == (int(want["root"]) - 1 + arena.base if want["root"] else 0)
                assert [int(d[k]) for k in ("votes", "total", "submit", "n_nodes", "first_node", "num_nodes")] == [int(want[k]) for k in ("votes", "total", "submit", "n_nodes", "first_node", "num_nodes")]
                assert np.array_equal(a, ca[offs[g]:offs[g + 1]]) and np.array_equal(b, co[offs[g]:offs[g + 1]])
    lp = np.where(z["sched_lp"] != 0, z["sched_lp"] - np.uint64(1) + np.uint64(arena.base), 0).astype(np.uint64)
    for sched in (0, 1, 2):
        for u in range(len(unit_offs) - 1):
            w0, w1 = int(unit_offs[u]), int(unit_offs[u + 1])
            assert port.schedule_pick(t, sched, int(lp[u]), lanes[32 * w0:32 * w1], stalled[w0:w1], front) == int(z["sched%d_pick" % sched][u])
    port2 = oracles.PortOracle(); port2.register(arena2); port2.form(budget)
    base = 0x7e0000000000
    for mode in (0, 1):
        r = port2.trace_remapped(mode, rays, base, stride, budget)
        p = "remap_m%d_" % mode
        assert np.array_equal(r["offsets"], z[p + "offsets"]) and np.array_equal(r["txns"], golden_util.replay_txns(z, p, base))
        assert np.array_equal(r["treelet_ids"], z[p + "tid"] + np.uint64(base))


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_table_events_match_reference():
    """Shader-table side effects with the reference's own Baseline tables (intersection_table.cc:165-187) in the loop:
    procedural leaves in both variants, any-hit calls + Hit_data for non-opaque rays in traceRay."""
    s = sc.Scene(1500, seed=8, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(1200, 12, 24, 16)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s); ref.form(512); port.register(s); port.form(512)
    ib, ab = ref.table_bases()
    seen = [0, 0]
    for mode in (0, 1):
        (ca, ea, ha), (cb, eb, hb) = ref.table_events(mode, rays), port.table_events(mode, rays, ib, ab)
        assert np.array_equal(ca, cb)
        for k in ("table", "shader_counter", "hit_group_index", "primitive_id", "instance_id", "tid", "store_addr", "store_size"):
            assert np.array_equal(ea[k], eb[k]), (mode, k)
        assert same_hits(ha, hb)
        seen[0] += int((ea["table"] == 0).sum()); seen[1] += int((ea["table"] == 1).sum())
        if mode == 1:
            assert not (ea["table"] == 1).any()          # traceRayWithTreelets has no any-hit path
    assert seen[0] >= 20 and seen[1] > 50 and int(ea["shader_counter"].max()) > 0


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_coalescing_table_matches_reference():
    """Function_Call_Coalescing intersection table (intersection_table.cc:43-98): the restatement replays the table over
    the table events; spliced into the plain trace it must give the transaction and store lists of the reference traversal
    run with its own Coalescing table (loads merged into the ray's list unless the address is already there)."""
    s = sc.Scene(1500, seed=8, n_blas=2, n_instances=5, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(1200, 12, 24, 16)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s); ref.form(512); port.register(s); port.form(512)
    stats = [0, 0, 0]
    # traceRay only: traceRayWithTreelets sends every record of the finished list through addrToTreeletID (:2256-2262), which
    # asserts on a table address (:470) -- the reference cannot run that variant with this table and procedural geometry
    for mode in (0,):
        to, tx, so, st, entry = ref.trace_coalescing(mode, rays)
        assert entry == helpers.COALESCING_ENTRY
        plain = port.trace(mode, rays)
        counts, ev, _ = port.table_events(mode, rays, 0, 0)
        eo = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        cev = port.coalescing_events(eo, ev)
        eto, etx, eso, est = helpers.coalescing_splice(plain["offsets"], plain["txns"], eo, ev["table"], ev["tid"], cev)
        assert np.array_equal(to, eto) and np.array_equal(so, eso), mode
        for k in ("address", "size", "type"):
            assert np.array_equal(tx[k], etx[k]), (mode, k)
            assert np.array_equal(st[k], est[k]), (mode, "store", k)
        t0 = ev["table"] == 0
        stats[0] += int((tx["type"] == 7).sum()); stats[1] += int(cev["appended"][t0].sum()); stats[2] += int((t0 & (cev["appended"] == 0)).sum())
    assert stats[0] > 40 and stats[1] > 10 and stats[2] > 10, stats      # loads merged, rows appended, rows shared between threads


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_unordered_child_bounds_match_reference():
    """A present child whose quantised lower bound exceeds its upper bound: the reference orders the slab planes with
    min / max, so the box is hit as if it were ordered; the restatement must agree (the CUDA fast path cannot -- K0 sends
    such an arena down the EXACT path, test_unordered_bounds_take_the_exact_path)."""
    a = helpers.kat_arena_unordered()
    rays = np.concatenate([helpers.kat_ray(0), helpers.kat_ray(1), sc.rays_random(200, seed=41)])
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    for budget in (256, 512):
        ref.register(a); ref.form(budget); port.register(a); port.form(budget)
        for mode in (0, 1):
            r, q = ref.trace(mode, rays), port.trace(mode, rays)
            assert np.array_equal(r["offsets"], q["offsets"]) and np.array_equal(r["txns"], q["txns"])
            assert np.array_equal(r["hits"]["prim"], q["hits"]["prim"])
            assert int(r["hits"]["hit"][1]) == 1 and int(r["offsets"][2] - r["offsets"][1]) == 9   # the opaque KAT ray still visits both quads and hits


def test_port_matches_coalescing_fixture():
    """The restatement's Coalescing-table replay against tests/golden/coalescing_proc1500.npz (recorded from the reference's
    traceRay with its own Coalescing table in the loop)."""
    z, arena, rays, tx, st = golden_util.load_coalescing()
    port = oracles.PortOracle(); port.register(arena); port.form(int(z["budget"]))
    plain = port.trace(0, rays)
    counts, ev, _ = port.table_events(0, rays, 0, 0)
    eo = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    cev = port.coalescing_events(eo, ev)
    eto, etx, eso, est = helpers.coalescing_splice(plain["offsets"], plain["txns"], eo, ev["table"], ev["tid"], cev)
    assert int(z["entry_size"]) == helpers.COALESCING_ENTRY
    assert np.array_equal(eto, z["txn_offsets"]) and np.array_equal(eso, z["store_offsets"])
    assert np.array_equal(etx, tx) and np.array_equal(est, st)


def test_port_matches_tables_fixture():
    """Shader-table events of the restatement against tests/golden/tables_proc1200.npz (recorded with the reference's own
    Baseline tables in the loop)."""
    z, arena, rays = golden_util.load_tables()
    port = oracles.PortOracle(); port.register(arena); port.form(int(z["budget"]))
    ib, ab = 0x6000000000, 0x6100000000
    for mode in (0, 1):
        counts, ev, ah = port.table_events(mode, rays, ib, ab)
        p = "m%d_" % mode
        assert np.array_equal(counts, z[p + "counts"])
        for k in ("table", "shader_counter", "hit_group_index", "primitive_id", "instance_id", "tid", "store_size"):
            assert np.array_equal(ev[k], z[p + k]), (mode, k)
        base = np.where(ev["table"] == 1, np.uint64(ab), np.uint64(ib))
        assert np.array_equal(ev["store_addr"] - base[:, None], z[p + "store_off"])
        want = z[p + "anyhit"].view(oracles.OHIT) if z[p + "anyhit"].dtype != oracles.OHIT else z[p + "anyhit"]
        assert same_hits(want, ah)


def _units(rng, n_rays, n_units):
    """Random RT units: 1-6 warps each, random (possibly repeated, possibly absent) rays per lane, some stalled."""
    offs = [0]; ids = []; st = []
    for _ in range(n_units):
        nw = int(rng.integers(1, 7))
        for _w in range(nw):
            lanes = rng.integers(0, n_rays, 32).astype(np.uint64)
            lanes[rng.random(32) < 0.2] = np.uint64(0xFFFFFFFFFFFFFFFF)
            ids.append(lanes); st.append(1 if rng.random() < 0.3 else 0)
        offs.append(offs[-1] + nw)
    return np.array(offs, np.uint64), np.concatenate(ids), np.array(st, np.uint8)


@pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_rt_unit_schedule_pick_matches_reference():
    """rt_unit::schedule_next_warp (shader.cc:4307-4392), -treelet_scheduler 0 / 1 (OMR) / 2 (PMR)."""
    s = sc.Scene(2500, seed=4, n_blas=2, n_instances=3)
    rays = sc.rays_primary(32, 16)
    ref, port = oracles.RefOracle(), oracles.PortOracle()
    ref.register(s); ref.form(512); port.register(s); port.form(512)
    t = port.trace(1, rays)
    rng = np.random.default_rng(11)
    offs, ids, st = _units(rng, len(rays), 60)
    roots = np.unique(t["treelet_ids"])
    picked = set()
    for u in range(len(offs) - 1):
        w0, w1 = int(offs[u]), int(offs[u + 1])
        front = rng.integers(0, 12, len(rays)).astype(np.uint32)
        for sched in (0, 1, 2):
            for lp in (0, int(roots[rng.integers(0, len(roots))]), int(t["treelet_ids"][int(t["offsets"][int(ids[32 * w0 + 3]) % len(rays)])])):
                a = ref.schedule_pick(t, sched, lp, ids[32 * w0:32 * w1], st[w0:w1], front)
                b = port.schedule_pick(t, sched, lp, ids[32 * w0:32 * w1], st[w0:w1], front)
                assert a == b, (u, sched, lp)
                picked.add((sched, a))
    assert len(picked) > 8


def test_port_parallel_equals_serial():
    s = sc.Scene(8000, seed=6, n_blas=2, n_instances=2)
    rays = helpers.mixed_rays(3000, 8)
    p1, p4 = oracles.PortOracle(), oracles.PortOracle()
    for p in (p1, p4):
        p.register(s); p.form(512)
    a, b = p1.trace(1, rays, nthreads=1), p4.trace(1, rays, nthreads=4)
    assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["txns"], b["txns"]) and same_hits(a["hits"], b["hits"])
    assert p1.counters() == p4.counters()


def test_properties():
    """Domain invariants of the reference semantics, checked on the port: TREELET and DFS agree on t for opaque
    rays; DFS reports the LAST accepted leaf while TREELET reports the closest."""
    s = sc.Scene(6000, seed=10)
    rays = sc.rays_primary(64, 48, flags=_abi.FLAG_OPAQUE)
    p = oracles.PortOracle(); p.register(s); p.form(512)
    d, t = p.trace(0, rays), p.trace(1, rays)
    assert np.array_equal(d["hits"]["hit"], t["hits"]["hit"])
    assert np.array_equal(d["hits"]["t"].view(np.uint32), t["hits"]["t"].view(np.uint32))
    assert d["hits"]["hit"].sum() > 100
    # every DFS trace is a superset-length walk: it never culls by the closest hit at leaves
    assert len(d["txns"]) >= len(t["txns"]) * 0.5

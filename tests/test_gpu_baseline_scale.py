"""GPU parity at the sizes BASELINE.json names: the CUDA path (through the C-ABI) against the reference's own code
(oracle/_ref; the C port when it did not travel) on the bench scene and on the incoherent-ray scene, plus the arena-size
edge the round-1 review found untested (a TLAS more than 512 MiB into its span)."""
import threading
import numpy as np
import pytest
from vsrt import scene as sc, _abi
import helpers
import oracles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    import vsrt.api as api
    return api


def best_oracle():
    return oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()


def strided(rays, n):
    return np.ascontiguousarray(rays[::max(1, len(rays) // n)][:n])


@pytest.fixture(scope="module")
def c2_scene():
    return sc.Scene(1_000_000, seed=0x5EED0001 + 1)      # the scene bench.py times (configs[1])


@pytest.mark.parametrize("budget", [512, 49152])
def test_c2_bench_scene_matches_reference(api, c2_scene, budget):
    """configs[1]: 1 M triangles, 1080p primary rays -- 32 K rays strided over the frame, both traversal variants, the shipped
    512 B budget and the 48 KB default; treelet tables of the whole 1 M-triangle BVH compared as well."""
    s = c2_scene
    rays = strided(sc.rays_primary(1920, 1080, flags=0), 32768)
    rays["ray_flags"][::3] = _abi.FLAG_OPAQUE
    rays["ray_flags"][::17] |= _abi.FLAG_TERMINATE_ON_FIRST_HIT
    orc = best_oracle()
    orc.register(s); orc.form(budget)
    ctx = api.Context(max_treelet_size=budget, device=0)
    try:
        ctx.register(s); ctx.form_treelets()
        helpers.assert_tables_equal(orc.tables(), ctx.tables(), "C2 budget %d" % budget)
        for mode in (_abi.MODE_TREELET, _abi.MODE_DFS):
            o = orc.trace(mode, rays); g = ctx.trace(mode, rays)
            helpers.assert_trace_equal(o, g, "C2 %s mode %d budget %d" % (orc.kind, mode, budget))
    finally:
        ctx.close()


def test_c3_incoherent_bounces_match_reference(api):
    """configs[2]: 2 M-triangle clustered scene, diffuse bounce rays generated from the CUDA path's own primary hits (the
    whole 1080p frame), bounces 1 and 2, 16 K strided rays each, treelet-ordered traversal."""
    s = sc.Scene(2_000_000, seed=0x5EED0001 + 2, kind=sc.CLUSTERED)
    orc = best_oracle()
    orc.register(s); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0)
    try:
        ctx.register(s); ctx.form_treelets()
        rays = sc.rays_primary(1920, 1080, flags=0)
        for bounce in (1, 2):
            full = ctx.trace(_abi.MODE_TREELET, rays, want_trace=False)
            rays = s.bounce(rays, full["hits"], 77, bounce - 1, 0)
            assert len(rays) > 50_000
            sample = strided(rays, 16384)
            o = orc.trace(_abi.MODE_TREELET, sample); g = ctx.trace(_abi.MODE_TREELET, sample)
            helpers.assert_trace_equal(o, g, "C3 bounce %d (%s)" % (bounce, orc.kind))
            assert len(o["txns"]) / len(sample) > 40      # these rays really are the long, incoherent ones
    finally:
        ctx.close()


def test_tlas_far_into_its_span(api):
    """A TLAS whose instance leaf sits more than 2^23 slots (512 MiB) past the first byte of its span: round 1 flagged such rays
    EF_UNSUPPORTED inside K1 and returned VSRT_OK with a truncated trace.  Instance references are now relative to the lowest
    instance leaf, so the layout (BLAS data merged below the TLAS, like a dump's .asback) traces normally."""
    gap = 545 * 1024 * 1024
    k = helpers.kat_arena()
    blas = bytes(k.bytes[320:576])                           # BLAS header, internal node, two quads
    data = np.zeros(gap + 320, np.uint8)
    data[:256] = np.frombuffer(blas, np.uint8)
    tl = bytearray(bytes(k.bytes[0:320]))                    # TLAS header, internal node, pad, instance leaf (128 B)
    leaf_off = gap + 192
    tl[192 + 64:192 + 72] = int((0 - leaf_off) % (1 << 64)).to_bytes(8, "little")   # BVHAddress: leaf -> BLAS header, relative
    data[gap:gap + 320] = np.frombuffer(bytes(tl), np.uint8)
    a = sc.Arena(data, tlas_offset=gap, blas=[(0, gap)])     # the BLAS buffer reaches up to the TLAS: one merged span
    del data
    rays = np.concatenate([helpers.kat_ray(0), helpers.kat_ray(1)])
    orc = best_oracle()
    orc.register(a); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0)
    try:
        ctx.register(a); ctx.form_treelets()
        helpers.assert_tables_equal(orc.tables(), ctx.tables(), "far TLAS")
        for mode in (0, 1):
            o = orc.trace(mode, rays); g = ctx.trace(mode, rays)
            helpers.assert_trace_equal(o, g, "far TLAS mode %d" % mode)
            assert len(g["txns"]) == 18 and g["hits"]["hit_geometry"][1] == 1      # (a ray without the Opaque flag never moves min_thit in traceRay)
    finally:
        ctx.close()


def test_stack_entries_limit_is_rejected(api):
    with pytest.raises(api.VsrtError) as e:
        api.Context(max_treelet_size=512, device=0, stack_entries=385)
    assert e.value.code == -1


def chain_arena(depth):
    """TLAS -> one instance -> a BLAS that is a chain: node i has a quad leaf (slot 0) and node i + 1 (slot 1); every box is
    the whole scene.  Later slots are popped first, so a ray through all the boxes keeps one pending entry per level."""
    a = helpers.mk_header(64, (-8,) * 3, (8,) * 3)
    a += helpers.mk_internal((-8, -8, -8), 2, (12, 12, 12), [6, 0, 0, 0, 0, 0], [(0, 0, 0)] * 6, [(1, 1, 1)] + [(0, 0, 0)] * 5)
    a += b"\0" * 64
    a += helpers.mk_instance(128, 7)
    a += helpers.mk_header(64, (-8,) * 3, (8,) * 3)
    for i in range(depth):
        last = i == depth - 1
        a += helpers.mk_internal((-8, -8, -8), 1, (12, 12, 12), [17, 17 if last else 1, 0, 0, 0, 0], [(0, 0, 0)] * 6, [(1, 1, 1), (1, 1, 1)] + [(0, 0, 0)] * 4)
        z = 1.0 + 0.01 * i
        a += helpers.mk_quad(100 + i, 3, [(-1, -1, z), (1, -1, z), (0, 1, z)])
    a += helpers.mk_quad(999, 3, [(-1, -1, 5.0), (1, -1, 5.0), (0, 1, 5.0)])
    return sc.Arena(a, 0, [(320, len(a) - 320)])


def test_stack_overflow_is_an_error_and_rolls_the_counters_back(api):
    """120 levels with a pending entry each overflow the default 96-entry stack in traceRay (the first box-hit internal child is
    followed at once, the quad beside it waits on the stack): VSRT_E_STACK_OVERFLOW, rayCount and the g_rt_* counters as before the
    call.  traceRayWithTreelets drains the quads of a treelet before it moves on and stays shallow.  With 192 entries the same
    rays match the reference in both variants."""
    a = chain_arena(120)
    away = helpers.kat_ray(0); away["direction"][0] = (0, 0, -1)
    ctx = api.Context(max_treelet_size=512, device=0)
    try:
        ctx.register(a); ctx.form_treelets()
        ctx.trace(1, away)
        before = ctx.counters()
        for _ in range(2):
            with pytest.raises(api.VsrtError) as e:
                ctx.trace(0, helpers.kat_ray(1))
            assert e.value.code == -7
            assert ctx.counters() == before
        ctx.trace(1, away)
        assert ctx.counters()["ray_count"] == before["ray_count"] + 1
    finally:
        ctx.close()
    rays = np.concatenate([helpers.kat_ray(1), helpers.kat_ray(0), away])
    orc = best_oracle()
    orc.register(a); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0, stack_entries=192)
    try:
        ctx.register(a); ctx.form_treelets()
        for mode in (0, 1):
            o = orc.trace(mode, rays); g = ctx.trace(mode, rays)
            helpers.assert_trace_equal(o, g, "chain mode %d" % mode)
    finally:
        ctx.close()


def test_two_gpu_reduce_inside_the_library(api):
    """vsrt_comm_init + vsrt_reduce_counters on two GPUs of this box (two contexts, one thread each): the global totals
    equal the single-context totals over the whole ray set, after several overlapping reduces."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from vsrt import shard
    s = sc.Scene(50000, seed=8, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    rays = np.concatenate([sc.rays_primary(256, 128, flags=0), sc.rays_random(32768, seed=4)])
    frames = 5
    uid = api.comm_unique_id()
    out, errs = {}, []

    def worker(rank):
        try:
            ctx = api.Context(max_treelet_size=512, device=rank)
            ctx.register(s); ctx.form_treelets()
            ctx.enable_node_histogram()
            ctx.comm_init(2, rank, uid)
            first, count = shard.shard_range(len(rays), 2, rank)
            for _ in range(frames):
                ctx.trace(1, rays[first:first + count], want_trace=False)
                ctx.reduce_counters()
            out[rank] = ctx.reduced() + (ctx.node_histogram(reduced=True),)
            ctx.close()
        except Exception as e:    # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ctx.form_treelets()
    ctx.enable_node_histogram()
    for _ in range(frames):
        ctx.trace(1, rays, want_trace=False)
    want, hist, nodes = ctx.counters(), ctx.treelet_histogram(), ctx.node_histogram()
    ctx.close()
    for rank in (0, 1):
        got, ghist, gnodes = out[rank]
        assert got == want, (rank, got, want)
        assert np.array_equal(ghist, hist) and np.array_equal(gnodes, nodes)

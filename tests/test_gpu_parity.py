"""GPU parity tests: the CUDA path, called through the C-ABI (libvsrt.so), against the oracle on the same seeded
inputs.  Oracle = oracle/_ref (the reference's own code) when it travelled with the snapshot, else the C port.
Bit-exact on node-visit sequence, record size/type, treelet ids, hit ids AND on t / barycentrics."""
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


def all_oracles():
    out = [oracles.PortOracle()]
    if oracles.have_ref():
        out.insert(0, oracles.RefOracle())
    return out


def run_case(api, arena, rays, budget, delta=0, blas_delta=None, modes=(0, 1), check_counters=True, stack_entries=96, skip_unknown_tids=False):
    for orc in all_oracles():
        orc.register(arena, delta, blas_delta)
        orc.form(budget)
        ctx = api.Context(max_treelet_size=budget, device=0, stack_entries=stack_entries)
        try:
            ctx.register(arena, delta, blas_delta)
            ctx.form_treelets()
            helpers.assert_tables_equal(orc.tables(), ctx.tables(), "%s budget %d" % (orc.kind, budget))
            for mode in modes:
                o = orc.trace(mode, rays)
                g = ctx.trace(mode, rays)
                helpers.assert_trace_equal(o, g, "%s mode %d budget %d delta %x" % (orc.kind, mode, budget, delta), skip_unknown_tids and mode == 0)
            if check_counters:
                co, cg = orc.counters(), ctx.counters()
                for i in range(9):
                    assert co["mem_access_type_%d" % i] == cg["mem_access_type_%d" % i], "type histogram bin %d" % i
                for ko, kg in helpers.COUNTER_MAP.items():
                    assert co[ko] == cg[kg], "counter %s: oracle %d cuda %d" % (ko, co[ko], cg[kg])
        finally:
            ctx.close()


def test_kat_vector(api):
    """SURVEY.md 8(c) worked vector, including the DFS/TREELET divergence when the two quads swap depth."""
    for z in ((1.0, 2.0), (2.0, 1.0)):
        for budget in (512, 256):
            a = helpers.kat_arena(*z)
            ctx = api.Context(max_treelet_size=budget, device=0)
            ctx.register(a)
            info = ctx.form_treelets()
            assert info.n_treelets == (1 if budget == 512 else 4)
            for mode in (0, 1):
                g = ctx.trace(mode, helpers.kat_ray(1 if mode == 0 else 0))
                rec = [(int(t["address"]) - a.base, int(t["size"]), int(t["type"])) for t in g["txns"]]
                last = 5 if (mode == 0 or z == (1.0, 2.0)) else 4
                assert rec == [(0, 64, 0), (64, 64, 1), (192, 128, 2), (320, 64, 0), (384, 64, 1), (512, 8, 3), (512, 64, 5), (448, 8, 3), (448, 64, last)]
                h = g["hits"][0]
                assert h["hit_geometry"] == 1 and h["geometry_index"] == 3 and h["instance_index"] == 7
                assert h["primitive_index"] == (101 if (mode == 1 and z == (2.0, 1.0)) else 100)
                assert h["world_min_thit"] == 6.0 and list(h["barycentric"]) == [0.25, 0.5, 0.25]
                tid = [int(x) - a.base for x in g["treelet_ids"]]
                assert tid == ([0] * 9 if budget == 512 else [0, 0, 192, 192, 192, 512, 512, 448, 448])
            ctx.close()
    run_case(api, helpers.kat_arena(), helpers.kat_ray(1), 256)


@pytest.mark.parametrize("ntri,nb,ni,flags,fan", [
    (300, 1, 1, 0, 6),
    (2000, 2, 3, sc.F_TRANSFORMS | sc.F_HOLES, 6),
    (5000, 3, 7, sc.F_TRANSFORMS, 4),
    (20000, 1, 2, sc.F_TRANSFORMS | sc.F_HOLES, 6),
])
@pytest.mark.parametrize("budget", [192, 512, 4096, 49152])
def test_random_scenes(api, ntri, nb, ni, flags, fan, budget):
    s = sc.Scene(ntri, seed=ntri, n_blas=nb, n_instances=ni, flags=flags, max_fanout=fan)
    run_case(api, s, helpers.mixed_rays(1500, ntri), budget)


@pytest.mark.parametrize("ntri,nb,ni,flags,budget,delta", [
    (2000, 2, 3, sc.F_TRANSFORMS | sc.F_HOLES, 512, 0),
    (20000, 1, 2, sc.F_TRANSFORMS | sc.F_HOLES, 49152, 0),
    (5000, 2, 4, sc.F_TRANSFORMS, 256, 0x100000),
])
def test_wavefront_variant(api, monkeypatch, ntri, nb, ni, flags, budget, delta):
    """The warp-wavefront formulation of K1 (traverse_wf.cu, VSRT_K1_WF=1): a pool of 64 rays per warp regrouped by
    phase every iteration.  Same records, treelet ids, hits and counters as the reference; also with non-finite rays
    (EXACT pass) mixed in."""
    monkeypatch.setenv("VSRT_K1_WF", "1")
    s = sc.Scene(ntri, seed=ntri + 1, n_blas=nb, n_instances=ni, flags=flags)
    rays = helpers.mixed_rays(3000, ntri + 5)
    rays["origin"][7::97, 1] = np.inf
    rays["direction"][11::89, 2] = np.nan
    run_case(api, s, rays, budget, delta=delta)


@pytest.mark.parametrize("ntri,nb,ni,flags,budget,delta", [
    (2000, 2, 3, sc.F_TRANSFORMS | sc.F_HOLES, 512, 0),
    (20000, 1, 2, sc.F_TRANSFORMS | sc.F_HOLES, 49152, 0),
    (30000, 3, 5, sc.F_TRANSFORMS, 4096, 0),
    (5000, 2, 4, sc.F_TRANSFORMS, 256, 0x100000),
    (1200, 1, 1, sc.F_PROCEDURAL, 1024, 0),
])
def test_treelet_binned_variant(api, monkeypatch, ntri, nb, ni, flags, budget, delta):
    """The treelet-binned wavefront formulation of K1 (traverse_tb.cu, VSRT_K1_TB=1): rounds of bin-by-next-treelet, TMA-staged
    shared treelets, drain.  traceRayWithTreelets records, treelet ids, hits and counters equal the reference's (traceRay runs
    through the lane-owned kernel either way); non-finite rays go through the EXACT pass; shared BLASes (instances > BLASes)
    exercise the "not stageable" fallback; a host->device offset exercises the :1752 quirk."""
    monkeypatch.setenv("VSRT_K1_TB", "1")
    s = sc.Scene(ntri, seed=ntri + 1, n_blas=nb, n_instances=ni, flags=flags)
    rays = helpers.mixed_rays(3000, ntri + 5)
    rays["origin"][7::97, 1] = np.inf
    rays["direction"][11::89, 2] = np.nan
    run_case(api, s, rays, budget, delta=delta)


def test_treelet_binned_variant_stages_treelets(api, monkeypatch):
    """At the reference's default 48 KB budget the top treelet is shared by every ray of the batch: the binned kernel must
    actually serve node visits from TMA-staged shared memory (the statistics say how many), and still match the default kernel."""
    s = sc.Scene(60000, seed=4)
    rays = np.concatenate([sc.rays_primary(128, 96, flags=0), sc.rays_random(8000, seed=9)])
    ctx = api.Context(max_treelet_size=49152, device=0)
    ctx.register(s); ctx.form_treelets()
    want = ctx.trace(1, rays)
    ctx.close()
    monkeypatch.setenv("VSRT_K1_TB", "1")
    ctx = api.Context(max_treelet_size=49152, device=0)
    ctx.register(s); ctx.form_treelets()
    got = ctx.trace(1, rays)
    st = ctx.tb_stats()
    ctx.close()
    for k in ("offsets", "txns", "treelet_ids", "hits"):
        assert np.array_equal(want[k], got[k]), k
    assert st["rounds"] > 2 and st["ctas_staged"] > 0 and st["visits_from_smem"] > 0
    assert st["visits_from_smem"] + st["visits_from_arena"] >= len(got["txns"]) // 3


def test_formation_store_grows_with_shared_blas(api):
    """Forty instances of one small BLAS under the 48 KB budget: every instance's treelet lists the whole BLAS again, so the
    lists hold many times more entries than the arena has nodes -- K0's compact list store starts at 1.25 entries per slot, has
    to grow and repeat a launch, and the tables must still equal the reference's (lists, de-duplication, highest-root-wins map)."""
    s = sc.Scene(300, seed=17, n_blas=1, n_instances=40, flags=sc.F_TRANSFORMS)
    for orc in all_oracles():
        orc.register(s); orc.form(49152)
        ctx = api.Context(max_treelet_size=49152, device=0)
        ctx.register(s); ti = ctx.form_treelets()
        to, tg = orc.tables(), ctx.tables()
        helpers.assert_tables_equal(to, tg, orc.kind)
        assert ti.n_list_entries > 1.25 * (s.size // 64) + 256        # the growth path really ran
        assert 0 < ti.scratch_bytes < (1 << 30)
        rays = helpers.mixed_rays(1500, 5)
        helpers.assert_trace_equal(orc.trace(1, rays), ctx.trace(1, rays), "shared BLAS, 48 KB")
        ctx.close()


def test_unordered_bounds_take_the_exact_path(api):
    """A present child with quantised lower > upper bound: the fast slab test reads near / far planes off the ray's
    direction signs and would miss the box, so K0 flags the arena and every ray runs the EXACT instantiation, which orders
    the planes with the reference's MIN / MAX.  Traces, hits, treelets and counters against the oracle."""
    rays = np.concatenate([helpers.kat_ray(0), helpers.kat_ray(1), sc.rays_random(300, seed=41)])
    for budget in (256, 512):
        run_case(api, helpers.kat_arena_unordered(), rays, budget)


def test_denormal_bound_scale_takes_the_exact_path(api):
    """Bound exponents below -118 make the scale 2^(e-8) a denormal float: the fast path builds the scale from the exponent
    byte alone, so K0 flags the arena and the EXACT instantiation (general ldexpf form) runs."""
    a = helpers.kat_arena()
    a.bytes[384 + 18:384 + 21] = (-120) & 0xff
    rays = np.concatenate([helpers.kat_ray(1), sc.rays_random(100, seed=43)])
    run_case(api, a, rays, 512)


def test_staging_overflow_regrows(api, monkeypatch):
    """A ray that outgrows its staging segment: the batch is redone with doubled segments (several times here: 8 records to
    start with) and the functional counters are rolled back in between -- traces, hits and counters must come out as if
    nothing had happened.  Procedural visits share the segment (their instance refs sit at its end)."""
    monkeypatch.setenv("VSRT_STAGE_CAP", "8")
    s = sc.Scene(2500, seed=31, n_blas=2, n_instances=4, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    run_case(api, s, helpers.mixed_rays(700, 31, 16, 12), 512)
    run_case(api, sc.Scene(1200, seed=32), helpers.mixed_rays(300, 32, 16, 12), 4096, modes=(0,))


@pytest.mark.parametrize("budget", [256, 1024])
def test_host_device_offset_quirk(api, budget):
    """Non-zero host->device offset: traceRayWithTreelets stores a HOST address in current_treelet_root
    (vulkan_ray_tracing.cc:1752), so after the first switch every child goes to the `other` stack."""
    s = sc.Scene(5000, seed=11, n_blas=2, n_instances=4, flags=sc.F_TRANSFORMS)
    run_case(api, s, helpers.mixed_rays(1000, 5), budget, delta=0x100000)


def test_non_uniform_blas_offsets(api):
    """BLAS buffers registered at their own device offsets: traceRay switches device_offset inside a BLAS (:2640),
    traceRayWithTreelets only for the BLAS header record (:1908-1913).  The device ranges are disjoint, as a bump
    allocator gives them (overlapping ranges would make the reference's address-keyed std::map collide)."""
    s = sc.Scene(3000, seed=21, n_blas=3, n_instances=5, flags=sc.F_TRANSFORMS)
    run_case(api, s, helpers.mixed_rays(800, 9), 512, delta=0x4000, blas_delta=[0x4000, 0x900000, 0x2000000], check_counters=False, skip_unknown_tids=True)


def test_clustered_scene_and_bounces(api):
    s = sc.Scene(30000, seed=5, kind=sc.CLUSTERED)
    prim = sc.rays_primary(96, 64, flags=_abi.FLAG_OPAQUE)
    orc = all_oracles()[0]
    orc.register(s); orc.form(2048)
    ctx = api.Context(max_treelet_size=2048, device=0)
    ctx.register(s); ctx.form_treelets()
    rays = prim
    for bounce in range(3):
        o = orc.trace(1, rays); g = ctx.trace(1, rays)
        helpers.assert_trace_equal(o, g, "bounce %d" % bounce)
        rays = s.bounce(rays, g["hits"], 77, bounce, 0)
        if len(rays) == 0:
            break
    ctx.close()


def test_node_visit_histogram(api):
    """The optional per-node visit histogram: records per node address over every batch since it was switched on, equal to a
    bincount of the returned trace; folding it through vsrt_node_treelet_table gives the per-treelet histogram."""
    s = sc.Scene(20000, seed=41, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    rays = helpers.mixed_rays(4000, 8)
    ctx = api.Context(max_treelet_size=1024, device=0)
    ctx.register(s); ctx.form_treelets()
    ctx.enable_node_histogram()
    want = np.zeros(s.size // 64, np.uint64)
    for mode in (1, 0, 1):
        g = ctx.trace(mode, rays)
        slots = ((g["txns"]["address"] - np.uint64(s.base)) >> np.uint64(6)).astype(np.int64)
        want += np.bincount(slots, minlength=len(want)).astype(np.uint64)
    got = ctx.node_histogram()
    assert np.array_equal(got, want)
    tab = ctx.node_treelet_table()
    th = np.bincount(tab[tab != 0xFFFFFFFF].astype(np.int64), weights=got[tab != 0xFFFFFFFF].astype(np.float64), minlength=ctx.treelet_info().n_treelets)
    assert np.array_equal(th.astype(np.uint64), ctx.treelet_histogram())
    ctx.reset_counters()
    assert ctx.node_histogram().sum() == 0
    ctx.close()


@pytest.mark.parametrize("budget", [512, 49152])
def test_treelet_histogram(api, budget):
    """The per-treelet visit histogram equals a bincount of the trace's treelet ids (metadata-index order), summed
    over batches and both variants; vsrt_reset_counters clears it."""
    s = sc.Scene(8000, seed=31, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    ctx = api.Context(max_treelet_size=budget, device=0)
    ctx.register(s); ctx.form_treelets()
    roots = ctx.tables()["roots"]
    expect = np.zeros(len(roots), np.uint64)
    for mode, rays in ((1, helpers.mixed_rays(3000, 1)), (0, sc.rays_primary(80, 60, flags=_abi.FLAG_OPAQUE)), (1, sc.rays_random(5000, seed=2))):
        g = ctx.trace(mode, rays)
        idx = np.searchsorted(roots, g["treelet_ids"])
        assert np.array_equal(roots[idx], g["treelet_ids"])
        expect += np.bincount(idx, minlength=len(roots)).astype(np.uint64)
    assert np.array_equal(ctx.treelet_histogram(), expect)
    ctx.reset_counters()
    assert ctx.treelet_histogram().sum() == 0
    ctx.close()


def test_nonfinite_rays_take_the_exact_path(api):
    """Rays with NaN/inf coordinates are deferred by the fast kernel (FMNMX slab test) to the EXACT kernel, which keeps
    the reference's ternary MIN/MAX; results must still match the oracle bit for bit."""
    s = sc.Scene(3000, seed=17, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    rays = helpers.mixed_rays(600, 4)
    rays["origin"][5::17, 0] = np.inf
    rays["direction"][3::19, 1] = np.nan
    rays["tmin"][7::23] = np.nan
    rays["tmax"][11::29] = np.inf
    rays["origin"][13::31, 2] = -np.inf
    with np.errstate(all="ignore"):
        run_case(api, s, rays, 512)
        run_case(api, s, rays, 4096, modes=(1,))


def test_empty_and_ragged_batches(api):
    s = sc.Scene(1000, seed=3)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ctx.form_treelets()
    g = ctx.trace(1, np.zeros(0, _abi.RAY))
    assert len(g["txns"]) == 0 and list(g["offsets"]) == [0]
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    for n in (1, 31, 33, 257):
        rays = sc.rays_random(n, seed=n)
        helpers.assert_trace_equal(orc.trace(1, rays), ctx.trace(1, rays), "n=%d" % n)
    # rays that miss the scene box entirely: exactly one record (the TLAS header) and no hit
    far = sc.rays_random(64, seed=1); far["origin"] += 100.0; far["direction"][:] = (1, 0, 0)
    g = ctx.trace(0, far)
    assert np.all(np.diff(g["offsets"]) == 1) and not g["hits"]["hit_geometry"].any()
    helpers.assert_trace_equal(orc.trace(0, far), g, "far rays")
    ctx.close()


def test_scan_tile_boundaries(api):
    """Batches whose ray counts sit on and around the tile size of the one-pass offset scan (8192 counts per tile; the look-back
    walks 32 tiles at a time): offsets, records and hits against the oracle."""
    s = sc.Scene(1500, seed=21, n_blas=2, n_instances=2)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ctx.form_treelets()
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    for n in (8191, 8192, 8193, 3 * 8192 + 1, 33 * 8192 + 7):
        rays = sc.rays_random(n, seed=n & 0xffff)
        helpers.assert_trace_equal(orc.trace(1, rays), ctx.trace(1, rays), "n=%d" % n)
    ctx.close()


def test_node_layout_is_picked_per_batch(api, monkeypatch):
    """Frame-sized batches queue both hot instantiations of K1 (traversal copy of the arena / Mesa layout) behind a device-side
    coherence sample; either layout, forced or picked, must give the oracle's traces -- for camera rays (picked: traversal copy)
    and for rays without any coherence (picked: Mesa layout)."""
    s = sc.Scene(6000, seed=23, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ctx.form_treelets()
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    cam = sc.rays_primary(320, 240)                # 76,800 rays: above the 65,536-ray threshold of the dispatch
    rnd = sc.rays_random(70000, seed=77)
    for rays, tag in ((cam, "camera"), (rnd, "random")):
        want = orc.trace(1, rays)
        for layout in (None, "0", "1"):
            if layout is None:
                monkeypatch.delenv("VSRT_K1_LAYOUT", raising=False)
            else:
                monkeypatch.setenv("VSRT_K1_LAYOUT", layout)
            helpers.assert_trace_equal(want, ctx.trace(1, rays), "%s rays, layout %s" % (tag, layout))
    monkeypatch.delenv("VSRT_K1_LAYOUT", raising=False)
    ctx.close()


def test_warp_call_and_queries(api):
    s = sc.Scene(4000, seed=8, n_blas=2, n_instances=2)
    orc = oracles.PortOracle(); orc.register(s); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0, treelet_based_traversal=1)
    ctx.register(s); ctx.form_treelets()
    rays = sc.rays_random(32, seed=4)
    mask = 0xF0F0FFFF
    w = ctx.trace_warp(rays, mask)
    act = [l for l in range(32) if mask >> l & 1]
    o = orc.trace(1, rays[act])
    assert list(w["counts"][act]) == list(np.diff(o["offsets"])) and w["counts"].sum() == len(o["txns"])
    assert np.array_equal(w["txns"], o["txns"])
    t = orc.tables()
    for node, root in list(zip(t["map_nodes"], t["map_roots"]))[::97]:
        assert ctx.addr_to_treelet(int(node)) == int(root)
    for i, r in enumerate(t["roots"][::53]):
        assert ctx.is_treelet_root(int(r)) and ctx.metadata_idx(int(r)) == i * 53
    assert not ctx.is_treelet_root(int(t["roots"][0]) + 8)
    with pytest.raises(api.VsrtError):
        ctx.addr_to_treelet(12345)
    ctx.close()


def test_error_paths(api):
    s = sc.Scene(500, seed=2)
    ctx = api.Context(max_treelet_size=128, device=0)
    ctx.register(s)
    with pytest.raises(api.VsrtError) as e:
        ctx.form_treelets()
    assert e.value.code == -8                      # budget below what the reference's asserts allow
    ctx.close()
    # BLAS never registered -> the reference asserts on blas_addr_map (:973); we return UNKNOWN_AS
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.alloc_tlas(s.tlas, s.size, s.tlas)
    with pytest.raises(api.VsrtError) as e:
        ctx.form_treelets()
    assert e.value.code == -5
    ctx.close()
    # PrimitiveIndex1Delta != 0 -> reference assert :2108
    a = helpers.kat_arena()
    a.bytes[448 + 12] = 1
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(a)
    with pytest.raises(api.VsrtError) as e:
        ctx.form_treelets()
    assert e.value.code == -6
    ctx.close()
    # unknown TLAS handle -> reference abort() (:1568)
    ctx = api.Context(device=0)
    ctx.register(s); ctx.tlas = s.tlas + 64
    with pytest.raises(api.VsrtError) as e:
        ctx.trace(0, sc.rays_random(4))
    assert e.value.code == -5
    ctx.close()


def test_remap_table(api):
    if not oracles.have_ref():
        pytest.skip("needs oracle/_ref")
    s = sc.Scene(3000, seed=13, n_blas=2, n_instances=3)
    ref = oracles.RefOracle()
    ref.register(s, remap=True, stride=256); ref.form(1024)
    base, o, m = ref.remap_table()
    ctx = api.Context(max_treelet_size=1024, device=0, treelet_remap_stride=256)
    ctx.register(s); ctx.form_treelets()
    go, gm = ctx.remap_table(base)
    assert np.array_equal(o, go) and np.array_equal(m, gm)
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_remapped_traces(api, mode):
    """-remap_to_treelet_layout 1: records and treelet ids in treelet-layout addresses, against the reference run with the
    option on (or the pinned port when oracle/_ref is absent)."""
    s = sc.Scene(3000, seed=13, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(2000, 21)
    if oracles.have_ref():
        ref = oracles.RefOracle(); ref.register(s, remap=True, stride=256); ref.form(1024)
        base = ref.remap_table()[0]
        o = ref.trace(mode, rays)
    else:
        base = 0x7f0000000000
        port = oracles.PortOracle(); port.register(s); port.form(1024)
        o = port.trace_remapped(mode, rays, base, 256, 1024)
    ctx = api.Context(max_treelet_size=1024, device=0, treelet_remap_stride=256, remap_to_treelet_layout=1)
    ctx.register(s); ctx.form_treelets()
    with pytest.raises(api.VsrtError):
        ctx.trace(mode, rays[:4])              # the layout base is an input (gpgpusim_malloc's answer in the reference)
    ctx.set_treelet_layout_base(base)
    g = ctx.trace(mode, rays)
    assert np.array_equal(o["offsets"], g["offsets"]) and np.array_equal(o["txns"], g["txns"])
    assert np.array_equal(o["treelet_ids"], g["treelet_ids"])
    ctx.close()


@pytest.mark.parametrize("budget", [512, 4096])
def test_sort_trace(api, budget):
    """rt_unit::sort_mem_accesses (shader.cc:3012-3089), both -sort_method values and both traversal variants, on a scene
    whose shared BLAS lists nodes in several treelets."""
    s = sc.Scene(3000, seed=13, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(1500, 23)
    orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
    orc.register(s); orc.form(budget)
    ctx = api.Context(max_treelet_size=budget, device=0); ctx.register(s); ctx.form_treelets()
    for mode in (0, 1):
        o = orc.trace(mode, rays)
        ctx.trace(mode, rays)
        for method in (1, 0, 1):                   # repeated calls always sort the ORIGINAL order
            want = orc.sort_trace(method, o)
            got, tids = ctx.sort_trace(method)
            assert np.array_equal(want, got), (mode, method)
            roots = orc.tables()
            idx = np.searchsorted(roots["map_nodes"], got["address"])
            assert np.array_equal(roots["map_roots"][idx], tids)        # the ids travel with their records
    ctx.close()


def test_prefetch_vote_and_chunks(api):
    """Treelet-prefetch vote of rt_unit::cycle (shader.cc:3419-3640) and the chunks it queues, per group of rays."""
    s = sc.Scene(3000, seed=13, n_blas=2, n_instances=3)
    rays = helpers.mixed_rays(6000, 29)
    orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
    orc.register(s); orc.form(1024)
    ctx = api.Context(max_treelet_size=1024, device=0); ctx.register(s); ctx.form_treelets()
    o = orc.trace(1, rays); ctx.trace(1, rays)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(rays)).astype(np.uint64)
    meta = (0x5000000000, (1024 // 64) * 4)
    for go, ids in ((np.arange(0, len(rays) + 1, 32), None), (np.array([0, 5, 5, 900, 6000]), None), (np.array([0, 64, 300, 5000]), perm)):
        for step in (0, 2, 7, 10 ** 6):
            front = rng.integers(0, step + 1, len(rays)).astype(np.uint32)
            for h, thr in ((0, 0.0), (1, 0.5), (2, 0.0), (3, 0.0)):
                for use_meta in (False, True):
                    dec = ctx.prefetch_vote(go, h, thr, ids, front, use_meta, meta[0])
                    offs, ca, co = ctx.prefetch_chunks(dec, h, use_meta, meta[0])
                    picks = range(len(go) - 1) if len(go) < 10 else rng.choice(len(go) - 1, 12, replace=False)
                    for g in picks:
                        gi = np.arange(go[g], go[g + 1]) if ids is None else ids[go[g]:go[g + 1]]
                        d, a, b = orc.prefetch_vote(o, gi, h, thr, front, meta if use_meta else None)
                        got = dec[g]
                        assert (int(got["treelet_root"]), int(got["votes"]), int(got["total"]), int(got["submit"]), int(got["n_nodes"]), int(got["first_node"]), int(got["num_nodes"])) == \
                               (int(d["root"]), int(d["votes"]), int(d["total"]), int(d["submit"]), int(d["n_nodes"]), int(d["first_node"]), int(d["num_nodes"])), (g, h, step)
                        assert np.array_equal(ca[offs[g]:offs[g + 1]], a) and np.array_equal(co[offs[g]:offs[g + 1]], b), (g, h, step)
    # one group larger than the shared-memory path
    big = rng.integers(0, len(rays), 9000).astype(np.uint64)
    dec = ctx.prefetch_vote(np.array([0, 9000]), 2, 0.0, big, None)
    d, a, b = orc.prefetch_vote(o, big, 2, 0.0, None, None)
    assert (int(dec[0]["treelet_root"]), int(dec[0]["votes"]), int(dec[0]["total"]), int(dec[0]["num_nodes"])) == (int(d["root"]), int(d["votes"]), int(d["total"]), int(d["num_nodes"]))
    ctx.close()


def ctx_offsets(ctx, n):
    """CSR trace offsets of the last batch, read back from the device through torch (the library keeps them resident)."""
    import torch
    r = ctx.device_results()

    class _D:
        __cuda_array_interface__ = {"shape": ((n + 1) * 8,), "typestr": "|u1", "data": (r.trace_offsets, False), "version": 3}
    return torch.as_tensor(_D(), device="cuda").cpu().numpy().view(np.uint64).astype(np.int64)


@pytest.mark.parametrize("wavefront", ["0", "1"])
def test_table_events(api, monkeypatch, wavefront):
    """Shader-table side effects (Baseline tables): procedural-leaf calls in both variants, any-hit calls + Hit_data of
    non-opaque rays in traceRay, against the reference with its own Baseline tables in the loop; shared thread indices
    (several threads of a CTA with the same tid.x) against the pinned restatement."""
    monkeypatch.setenv("VSRT_K1_WF", wavefront)
    s = sc.Scene(2500, seed=18, n_blas=2, n_instances=4, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(1000, 33, 16, 12)
    port = oracles.PortOracle(); port.register(s); port.form(512)
    ref = None
    if oracles.have_ref():
        ref = oracles.RefOracle(); ref.register(s); ref.form(512)
    ib, ab = ref.table_bases() if ref else (0x6000000000, 0x6100000000)
    ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
    n_ev = [0, 0]
    for mode in (0, 1):
        ctx.trace(mode, rays)
        for tid_x in (None, (np.arange(len(rays)) % 8).astype(np.uint8)):
            co, eo, ho = (ref if (ref and tid_x is None) else port).table_events(mode, rays, *(() if (ref and tid_x is None) else (ib, ab, tid_x)))
            offs, ev, ah = ctx.table_events(tid_x)
            assert np.array_equal(np.diff(offs).astype(np.uint32), co), (mode, "counts")
            for k in ("table", "shader_counter", "hit_group_index", "primitive_id", "instance_id", "tid"):
                assert np.array_equal(eo[k], ev[k]), (mode, k)
            st = ctx.table_event_stores(ev, (ib, ab))
            assert np.array_equal(st["address"], eo["store_addr"]) and np.array_equal(st["size"], eo["store_size"])
            anyh = ev["table"] == 1
            assert np.array_equal(ho["t"][anyh].view(np.uint32), ah["world_min_thit"][anyh].view(np.uint32))
            assert np.array_equal(ho["prim"][anyh], ah["primitive_index"][anyh]) and np.array_equal(ho["instance_id"][anyh], ah["instance_index"][anyh])
            assert np.array_equal(ho["bary"][anyh].view(np.uint32), ah["barycentric"][anyh].view(np.uint32))
            assert np.array_equal(ho["point"][anyh].view(np.uint32), ah["intersection_point"][anyh].view(np.uint32))
            # the event's record is the PROCEDURAL_LEAF (6) / QUAD_LEAF_HIT (5) record of its ray
            t = ctx.fetch_trace()[0]
            toffs = np.asarray(ctx_offsets(ctx, len(rays)))
            ray_of = np.repeat(np.arange(len(rays)), np.diff(offs).astype(np.int64))
            assert np.array_equal(t["type"][toffs[ray_of] + ev["record"]], np.where(ev["table"] == 0, 6, 5))
            n_ev[0] += int((ev["table"] == 0).sum()); n_ev[1] += int(anyh.sum())
    assert n_ev[0] > 50 and n_ev[1] > 100
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_packed_trace(api, mode):
    """vsrt_trace_rays_packed / vsrt_trace_fetch_packed: 8 bytes per record instead of 24; expanded with vsrt_unpack_txn and the
    treelet table they must give exactly the records and treelet ids of vsrt_trace_rays (which the other tests pin to the
    oracle), also for an arena of several host spans and a host->device offset."""
    s = sc.Scene(3000, seed=21, n_blas=3, n_instances=4, flags=sc.F_TRANSFORMS)
    rays = helpers.mixed_rays(1500, 21, 24, 16)
    ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s, delta=0x1000000); ctx.form_treelets()
    g = ctx.trace(mode, rays)
    rec, tix = ctx.fetch_packed()
    assert len(rec) == len(g["txns"])
    assert np.array_equal(ctx.unpack(rec), g["txns"])
    roots = ctx.tables()["roots"]
    ids = np.where(tix == 0xFFFFFFFF, np.uint64(0xFFFFFFFFFFFFFFFF), roots[np.minimum(tix, len(roots) - 1)])
    assert np.array_equal(ids, g["treelet_ids"])
    # one call, host buffers
    n = len(rays)
    hits = np.zeros(n, _abi.HIT); offs = np.zeros(n + 1, np.uint64); rec2 = np.zeros(len(rec), np.uint32); tix2 = np.zeros(len(rec), np.uint32)
    r = np.ascontiguousarray(rays, _abi.RAY)
    got = ctx.trace_packed_into(mode, n, r.ctypes.data, hits.ctypes.data, offs.ctypes.data, rec2.ctypes.data, len(rec2), tix2.ctypes.data)
    assert got == len(rec) and np.array_equal(rec2, rec) and np.array_equal(tix2, tix) and np.array_equal(offs, g["offsets"])
    assert np.array_equal(hits["primitive_index"], g["hits"]["primitive_index"])
    ctx.close()
    # non-uniform BLAS offsets: the two address conventions of the reference differ, the packed form cannot express them
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s, delta=0x4000, blas_delta=[0x4000, 0x900000, 0x2000000]); ctx.form_treelets()
    ctx.trace(mode, rays[:64])
    with pytest.raises(api.VsrtError) as e:
        ctx.fetch_packed()
    assert e.value.code == -9
    ctx.close()


def test_packed_pipeline_and_treelet_table(api, monkeypatch):
    """The lean host form: packed records without the treelet-index stream, traced in chunks with overlapped copies
    (VSRT_PIPELINE_CHUNK small enough to force several chunks, ragged last one).  Same hits, offsets and records as the
    one-batch call; vsrt_node_treelet_table()[record >> 3] reproduces the treelet index of every record; the counters count
    every ray once; a later vsrt_trace_rays works as before."""
    s = sc.Scene(20000, seed=13, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    rays = helpers.mixed_rays(9000, 3)
    ctx = api.Context(max_treelet_size=512, device=0)
    ctx.register(s); ctx.form_treelets()
    monkeypatch.setenv("VSRT_PIPELINE_CHUNK", "0")
    h0, o0, r0, t0 = ctx.trace_packed(1, rays, want_index=True)
    c0 = ctx.counters()
    full = ctx.trace(1, rays)
    ctx.reset_counters()
    monkeypatch.setenv("VSRT_PIPELINE_CHUNK", "2048")
    h1, o1, r1, _ = ctx.trace_packed(1, rays)
    c1 = ctx.counters()
    assert np.array_equal(h0, h1) and np.array_equal(o0, o1) and np.array_equal(r0, r1)
    assert np.array_equal(o0, full["offsets"]) and np.array_equal(ctx.unpack(r1), full["txns"])
    tab = ctx.node_treelet_table()
    assert np.array_equal(tab[r1 >> 3], t0)
    assert c1 == c0                                   # both ways trace the batch twice (size query, then the data): same counters
    # after a pipelined call the context holds the whole frame, like after a single batch: the replay helpers work on it
    txn_dev, tid_dev = ctx.fetch_trace()
    assert np.array_equal(txn_dev, full["txns"]) and np.array_equal(tid_dev, full["treelet_ids"])
    # the full host form, pipelined: records copied window by window, treelet ids derived on the host from the slot -> root table
    # (VSRT_HOST_EXPAND=0), or -- the default -- windows delivered as 4-byte packed records and expanded to records + ids by host threads
    for expand in ("0", "1"):
        monkeypatch.setenv("VSRT_HOST_EXPAND", expand)
        again = ctx.trace(1, rays, capacity=len(full["txns"]))
        assert np.array_equal(again["offsets"], full["offsets"]) and np.array_equal(again["hits"], full["hits"]), expand
        assert np.array_equal(again["txns"], full["txns"]) and np.array_equal(again["treelet_ids"], full["treelet_ids"]), expand
        txn_dev, tid_dev = ctx.fetch_trace()          # the device-side frame is complete either way
        assert np.array_equal(txn_dev, full["txns"]) and np.array_equal(tid_dev, full["treelet_ids"]), expand
    monkeypatch.delenv("VSRT_HOST_EXPAND")
    sorted_txns, _ = ctx.sort_trace(1)
    monkeypatch.setenv("VSRT_PIPELINE_CHUNK", "0")
    ctx.trace(1, rays)
    want_sorted, _ = ctx.sort_trace(1)
    assert np.array_equal(sorted_txns, want_sorted)
    ctx.close()


def test_coalescing_table(api):
    """Function_Call_Coalescing intersection table: vsrt_coalescing_events over the CUDA path's own table events against the
    pinned restatement, and the spliced transaction / store lists against the reference traversal run with its own
    Coalescing table (oracle/_ref) where that is available."""
    s = sc.Scene(2500, seed=18, n_blas=2, n_instances=6, flags=sc.F_TRANSFORMS | sc.F_PROCEDURAL)
    rays = helpers.mixed_rays(1000, 33, 16, 12)
    port = oracles.PortOracle(); port.register(s); port.form(512)
    ref = None
    if oracles.have_ref():
        ref = oracles.RefOracle(); ref.register(s); ref.form(512)
    ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
    n_loads = 0
    for mode in (0, 1):
        g = ctx.trace(mode, rays)
        offs, ev, _ = ctx.table_events()
        cev = ctx.coalescing_events(offs, ev)
        pev = np.zeros(len(ev), oracles.TEV)
        for k in ("table", "hit_group_index", "tid"):
            pev[k] = ev[k]
        want = port.coalescing_events(offs, pev)
        for k in ("row", "appended", "n_loads", "first_new_load"):
            assert np.array_equal(cev[k], want[k]), (mode, k)
        to, tx, so, st = ctx.coalescing_trace(g["offsets"], g["txns"], offs, ev, cev, 1 << 63)
        eto, etx, eso, est = helpers.coalescing_splice(g["offsets"], g["txns"], offs, ev["table"], ev["tid"], want)
        assert np.array_equal(to, eto) and np.array_equal(so, eso)
        assert np.array_equal(tx, etx) and np.array_equal(st, est)
        if ref and mode == 0:   # the reference's treelet variant asserts in addrToTreeletID (:470) on the first table address in a list
            rto, rtx, rso, rst, _ = ref.trace_coalescing(mode, rays)
            assert np.array_equal(rto, to) and np.array_equal(rso, so), mode
            for k in ("address", "size", "type"):
                assert np.array_equal(rtx[k], tx[k]), (mode, k)
                assert np.array_equal(rst[k], st[k]), (mode, "store", k)
        n_loads += int((tx["type"] == 7).sum())
    assert n_loads > 40
    # more than INTERSECTION_TABLE_MAX_LENGTH rows in one CTA: the reference would write past its allocation
    many = np.zeros(101, _abi.TEV); many["hit_group_index"] = np.arange(101)
    with pytest.raises(api.VsrtError) as e:
        ctx.coalescing_events(np.array([0, 101], np.uint64), many)
    assert e.value.code == -9
    ctx.close()


def test_schedule_pick(api):
    """rt_unit::schedule_next_warp (shader.cc:4307-4392) for many units at once."""
    from test_oracle import _units
    s = sc.Scene(3000, seed=13, n_blas=2, n_instances=3)
    rays = sc.rays_primary(64, 32)
    orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
    orc.register(s); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0); ctx.register(s); ctx.form_treelets()
    o = orc.trace(1, rays); ctx.trace(1, rays)
    rng = np.random.default_rng(17)
    offs, ids, st = _units(rng, len(rays), 200)
    roots = np.unique(o["treelet_ids"])
    front = rng.integers(0, 12, len(rays)).astype(np.uint32)
    lp = roots[rng.integers(0, min(len(roots), 6), len(offs) - 1)]          # popular treelets near the top of the tree
    lp[::7] = 0
    for sched in (0, 1, 2):
        got = ctx.schedule_pick(sched, offs, ids, st, lp, front)
        for u in range(len(offs) - 1):
            w0, w1 = int(offs[u]), int(offs[u + 1])
            want = orc.schedule_pick(o, sched, int(lp[u]), ids[32 * w0:32 * w1], st[w0:w1], front)
            assert (int(got[u]) - w0 if got[u] >= 0 else -1) == want, (sched, u)
    ctx.close()


@pytest.mark.parametrize("order", ["2"])
def test_ray_order_never_changes_results(api, monkeypatch, order):
    """K1 may pick the rays of a batch up in sorted order (rayorder.cu: origin cell + direction key, radix sort); every output is
    indexed by the ray's own position, so traces, hits and counters are those of the input order and of the reference."""
    monkeypatch.setenv("VSRT_RAY_ORDER_MIN", "1")
    s = sc.Scene(20000, seed=31, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    batches = [helpers.mixed_rays(6000, 17), sc.rays_primary(80, 60, flags=1), sc.rays_random(5000, seed=3)[:4097]]
    orc = all_oracles()[0]
    orc.register(s); orc.form(512)
    res = {}
    for ro in ("1", order):
        monkeypatch.setenv("VSRT_RAY_ORDER", ro)
        ctx = api.Context(max_treelet_size=512, device=0)
        ctx.register(s); ctx.form_treelets()
        res[ro] = [ctx.trace(mode, b) for b in batches for mode in (0, 1)]
        if ro == order:
            r = ctx.device_results()
            assert r.order_ms > 0.0
        res[ro + "c"] = ctx.counters()
        ctx.close()
    assert res["1c"] == res[order + "c"]
    i = 0
    for b in batches:
        for mode in (0, 1):
            o = orc.trace(mode, b)
            helpers.assert_trace_equal(o, res[order][i], "sorted batch %d mode %d" % (i // 2, mode))
            for k in ("offsets", "txns", "treelet_ids", "hits"):
                assert np.array_equal(res["1"][i][k], res[order][i][k])
            i += 1


def test_large_scene_properties(api):
    """Full-size style check through size-independent properties (no oracle): both variants agree on hit t
    for opaque closest-hit rays; per-ray records start with the TLAS header; counters equal the trace."""
    s = sc.Scene(200000, seed=99)
    rays = sc.rays_primary(320, 240, flags=_abi.FLAG_OPAQUE)
    ctx = api.Context(max_treelet_size=4096, device=0)
    ctx.register(s); ctx.form_treelets()
    a = ctx.trace(0, rays); b = ctx.trace(1, rays)
    assert np.array_equal(a["hits"]["hit_geometry"], b["hits"]["hit_geometry"])
    assert np.array_equal(a["hits"]["world_min_thit"].view(np.uint32), b["hits"]["world_min_thit"].view(np.uint32))
    for g in (a, b):
        first = g["txns"][g["offsets"][:-1].astype(np.int64)]
        assert np.all(first["address"] == s.tlas) and np.all(first["type"] == 0)
    c = ctx.counters()
    total = sum(c["mem_access_type_%d" % i] for i in range(9))
    assert total == len(a["txns"]) + len(b["txns"])
    assert c["accessed_data_size"] == int(a["txns"]["size"].sum()) + int(b["txns"]["size"].sum())
    assert c["ray_count"] == 2 * len(rays)
    # every treelet id is the root of the treelet owning the record's node
    t = ctx.tables()
    idx = np.searchsorted(t["map_nodes"], b["txns"]["address"])
    assert np.array_equal(t["map_roots"][idx], b["treelet_ids"])
    ctx.close()


import golden_util  # noqa: E402


@pytest.mark.parametrize("path", golden_util.fixtures(), ids=lambda p: p.split("golden_")[-1][:-4])
def test_cuda_matches_golden(api, path):
    """CUDA path against the fixtures recorded from the reference's own code (tests/golden/make_golden.py)."""
    z, arena, rays = golden_util.load(path)
    delta = int(z["delta"])
    for b in (int(x) for x in z["budgets"]):
        ctx = api.Context(max_treelet_size=b, device=0)
        ctx.register(arena, delta); ctx.form_treelets()
        helpers.assert_tables_equal(golden_util.expected_tables(z, b, arena.base), ctx.tables(), "golden budget %d" % b)
        for mode in (0, 1):
            helpers.assert_trace_equal(golden_util.expected_trace(z, b, mode, arena.base), ctx.trace(mode, rays), "golden b%d m%d" % (b, mode))
        ctx.close()


def test_cuda_matches_replay_fixture(api):
    """rt_unit helpers and remapped traces of the CUDA path against tests/golden/replay_inst1500.npz (recorded from the
    reference's own rt_unit bodies and its -remap_to_treelet_layout traversal)."""
    z, arena, arena2, rays = golden_util.load_replay()
    budget, stride = int(z["budget"]), int(z["stride"])
    go, front, unit_offs, lanes, stalled = golden_util.replay_groups(len(rays))
    ctx = api.Context(max_treelet_size=budget, device=0); ctx.register(arena); ctx.form_treelets()
    for mode in (0, 1):
        ctx.trace(mode, rays)
        for method in (0, 1):
            got, _ = ctx.sort_trace(method)
            assert np.array_equal(got, golden_util.replay_txns(z, "m%d_s%d_" % (mode, method), arena.base)), (mode, method)
    ctx.trace(1, rays)
    for h, thr in ((0, 0.0), (1, 0.4), (2, 0.0), (3, 0.0)):
        for use_meta in (0, 1):
            p = "h%d_meta%d_" % (h, use_meta)
            dec = ctx.prefetch_vote(go, h, thr, None, front, bool(use_meta), golden_util.META_BASE)
            offs, ca, co = ctx.prefetch_chunks(dec, h, bool(use_meta), golden_util.META_BASE)
            want = z[p + "dec"]
            assert np.array_equal(dec["treelet_root"], np.where(want["root"] != 0, want["root"] - np.uint64(1) + np.uint64(arena.base), 0).astype(np.uint64))
            for k in ("votes", "total", "submit", "n_nodes", "first_node", "num_nodes"):
                assert np.array_equal(dec[k], want[k]), (h, use_meta, k)
            woffs, wca, wco = golden_util.replay_chunks(z, p, arena.base)
            assert np.array_equal(offs, woffs) and np.array_equal(ca, wca) and np.array_equal(co, wco)
    lp = np.where(z["sched_lp"] != 0, z["sched_lp"] - np.uint64(1) + np.uint64(arena.base), 0).astype(np.uint64)
    for sched in (0, 1, 2):
        got = ctx.schedule_pick(sched, unit_offs, lanes, stalled, lp, front)
        want = z["sched%d_pick" % sched]
        assert np.array_equal(np.where(got >= 0, got - unit_offs[:-1].astype(np.int64), -1), want), sched
    ctx.close()
    base = 0x7e0000000000
    ctx = api.Context(max_treelet_size=budget, device=0, treelet_remap_stride=stride, remap_to_treelet_layout=1)
    ctx.register(arena2); ctx.form_treelets(); ctx.set_treelet_layout_base(base)
    for mode in (0, 1):
        g = ctx.trace(mode, rays)
        p = "remap_m%d_" % mode
        assert np.array_equal(g["offsets"], z[p + "offsets"]) and np.array_equal(g["txns"], golden_util.replay_txns(z, p, base))
        assert np.array_equal(g["treelet_ids"], z[p + "tid"] + np.uint64(base))
    ctx.close()


def test_cuda_matches_tables_fixture(api):
    """vsrt_table_events against tests/golden/tables_proc1200.npz (the reference's Baseline tables in the loop)."""
    z, arena, rays = golden_util.load_tables()
    ctx = api.Context(max_treelet_size=int(z["budget"]), device=0); ctx.register(arena); ctx.form_treelets()
    ib, ab = 0x6000000000, 0x6100000000
    for mode in (0, 1):
        ctx.trace(mode, rays)
        offs, ev, ah = ctx.table_events()
        p = "m%d_" % mode
        assert np.array_equal(np.diff(offs).astype(np.uint32), z[p + "counts"])
        for k in ("table", "shader_counter", "hit_group_index", "primitive_id", "instance_id", "tid"):
            assert np.array_equal(ev[k], z[p + k]), (mode, k)
        st = ctx.table_event_stores(ev, (ib, ab))
        base = np.where(ev["table"] == 1, np.uint64(ab), np.uint64(ib))
        assert np.array_equal(st["address"] - base[:, None], z[p + "store_off"]) and np.array_equal(st["size"], z[p + "store_size"])
        want = z[p + "anyhit"]
        anyh = ev["table"] == 1
        assert np.array_equal(want["t"][anyh].view(np.uint32), ah["world_min_thit"][anyh].view(np.uint32))
        assert np.array_equal(want["bary"][anyh].view(np.uint32), ah["barycentric"][anyh].view(np.uint32))
        assert np.array_equal(want["prim"][anyh], ah["primitive_index"][anyh])
    ctx.close()


def test_cuda_matches_coalescing_fixture(api):
    """vsrt_coalescing_events spliced into the CUDA trace against tests/golden/coalescing_proc1500.npz (the reference's
    traceRay with its own Coalescing table in the loop)."""
    z, arena, rays, tx, st = golden_util.load_coalescing()
    ctx = api.Context(max_treelet_size=int(z["budget"]), device=0); ctx.register(arena); ctx.form_treelets()
    g = ctx.trace(0, rays)
    offs, ev, _ = ctx.table_events()
    cev = ctx.coalescing_events(offs, ev)
    to, gtx, so, gst = ctx.coalescing_trace(g["offsets"], g["txns"], offs, ev, cev, 1 << 63)
    assert np.array_equal(to, z["txn_offsets"]) and np.array_equal(so, z["store_offsets"])
    assert np.array_equal(gtx, tx) and np.array_equal(gst, st)
    ctx.close()


"""Property tests (hypothesis) of the host-side pieces that have a closed-form contract: ray-id sharding and the
Function_Call_Coalescing table replay of the oracle restatement (the CUDA kernel is compared with it in test_gpu_parity)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracles
from vsrt import shard, _abi


@settings(max_examples=200, deadline=None)
@given(total=st.integers(0, 1 << 27), world=st.integers(1, 16))
def test_shard_ranges_partition_the_frame(total, world):
    """Contiguous blocks in rank order that cover [0, total) exactly once; every block starts on a warp boundary (a warp's
    lanes never straddle two ranks) and no two blocks differ by more than one warp."""
    pos = 0
    sizes = []
    for r in range(world):
        first, count = shard.shard_range(total, world, r)
        assert first == pos or count == 0
        assert first % shard.WARP == 0 or count == 0
        pos = first + count if count else pos
        sizes.append(count)
    assert pos == total and sum(sizes) == total
    full = [c for c in sizes if c % shard.WARP == 0]
    if len(full) > 1:
        assert max(full) - min(full) <= shard.WARP


def _events(rng, n_rays, max_ev, n_groups):
    counts = rng.integers(0, max_ev + 1, n_rays)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    ev = np.zeros(int(offs[-1]), oracles.TEV)
    ev["table"] = (rng.random(len(ev)) < 0.2).astype(np.uint32)
    ev["hit_group_index"] = rng.integers(0, n_groups, len(ev))
    ray_of = np.repeat(np.arange(n_rays), counts)
    ev["tid"] = ray_of % 32
    return offs, ev, ray_of


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 1 << 30), n_rays=st.integers(1, 200), max_ev=st.integers(0, 6), n_groups=st.integers(1, 12))
def test_coalescing_replay_invariants(seed, n_rays, max_ev, n_groups):
    """What intersection_table.cc:43-98 implies for any event stream: inside a CTA (32 rays) a (row, thread) pair is claimed at
    most once; a row holds one hit group; rows are appended in order; a claim of row i looked at i + 1 rows, an append at all
    existing rows; the loads new to a ray are exactly the rows beyond the most it has looked at before."""
    rng = np.random.default_rng(seed)
    offs, ev, ray_of = _events(rng, n_rays, max_ev, n_groups)
    port = oracles.PortOracle()
    cev = port.coalescing_events(offs, ev)
    assert np.all(cev["row"][ev["table"] == 1] == 0) and np.all(cev["n_loads"][ev["table"] == 1] == 0)
    for g0 in range(0, n_rays, 32):
        sel = np.flatnonzero((ray_of >= g0) & (ray_of < g0 + 32) & (ev["table"] == 0))
        claimed, row_key, n_rows = set(), {}, 0
        seen = {}
        for k in sel:
            row, app, nl, fn = (int(cev[k][f]) for f in ("row", "appended", "n_loads", "first_new_load"))
            tid, key = int(ev[k]["tid"]), int(ev[k]["hit_group_index"])
            assert (row, tid) not in claimed
            claimed.add((row, tid))
            if app:
                assert row == n_rows and nl == n_rows
                # nothing earlier could have taken it: every existing row of this group already holds this thread
                assert all((i, tid) in claimed for i in range(n_rows) if row_key[i] == key)
                row_key[row] = key; n_rows += 1
            else:
                assert row < n_rows and row_key[row] == key and nl == row + 1
                assert all((i, tid) in claimed for i in range(row) if row_key[i] == key)
            r = int(ray_of[k])
            assert fn == min(seen.get(r, 0), nl)
            seen[r] = max(seen.get(r, 0), nl)

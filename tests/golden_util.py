"""Replay helpers for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference's own code)."""
import glob
import os
import numpy as np
from vsrt import scene as sc, _abi
import oracles

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "golden_*.npz")))


def load(path):
    z = np.load(path)
    arena = sc.Arena(z["arena"], int(z["tlas_offset"]), [(int(o), int(s)) for o, s in z["blas"]])
    return z, arena, z["rays"].view(_abi.RAY) if z["rays"].dtype != _abi.RAY else z["rays"]


def expected_tables(z, b, base):
    t = {}
    for k in ("roots", "node_addr", "map_nodes", "map_roots"):
        t[k] = z["b%d_%s" % (b, k)] + np.uint64(base)
    for k in ("counts", "meta_idx", "node_size"):
        t[k] = z["b%d_%s" % (b, k)]
    return t


def expected_trace(z, b, mode, base):
    p = "b%d_m%d_" % (b, mode)
    n = len(z[p + "addr"])
    txns = np.zeros(n, _abi.TXN)
    txns["address"] = z[p + "addr"] + np.uint64(base); txns["size"] = z[p + "size"]; txns["type"] = z[p + "type"]
    return {"offsets": z[p + "offsets"], "txns": txns, "treelet_ids": z[p + "tid"] + np.uint64(base),
            "hits": z[p + "hits"].view(oracles.OHIT) if z[p + "hits"].dtype != oracles.OHIT else z[p + "hits"]}


# ---- replay_*.npz: rt_unit helpers + remapped traversal recorded from the reference (make_golden.main_replay) ----
REPLAY = os.path.join(GOLDEN_DIR, "replay_inst1500.npz")
META_BASE = 0x5000000000


def replay_groups(n_rays):
    # the group layout is a pure function of n_rays; re-derive it here instead of importing the generator (which builds the reference)
    rng = np.random.default_rng(77)
    go = np.array([0, 32, 64, 200, n_rays], np.uint64)
    front = rng.integers(0, 9, n_rays).astype(np.uint32)
    unit_offs = np.array([0, 3, 4, 9], np.uint64)
    lanes = rng.integers(0, n_rays, 9 * 32).astype(np.uint64)
    lanes[rng.random(len(lanes)) < 0.15] = np.uint64(0xFFFFFFFFFFFFFFFF)
    stalled = np.array([0, 1, 0, 0, 1, 0, 0, 0, 1], np.uint8)
    return go, front, unit_offs, lanes, stalled


def load_replay():
    z = np.load(REPLAY)
    arena = sc.Arena(z["arena"], int(z["tlas_offset"]), [(int(o), int(s)) for o, s in z["blas"]])
    arena2 = sc.Arena(z["remap_arena"], int(z["remap_tlas_offset"]), [(int(o), int(s)) for o, s in z["remap_blas"]])
    rays = z["rays"].view(_abi.RAY) if z["rays"].dtype != _abi.RAY else z["rays"]
    return z, arena, arena2, rays


def replay_txns(z, prefix, base):
    n = len(z[prefix + "addr"])
    t = np.zeros(n, _abi.TXN)
    t["address"] = z[prefix + "addr"] + np.uint64(base); t["size"] = z[prefix + "size"]; t["type"] = z[prefix + "type"]
    return t


def replay_chunks(z, prefix, base):
    """(offsets, chunk addresses, owners) with node chunks rebased to `base` (metadata rows are absolute)."""
    ca, co = z[prefix + "chunk_addr"].copy(), z[prefix + "chunk_owner"].copy()
    node = co < np.uint64(META_BASE)
    ca[node] += np.uint64(base); co[node] += np.uint64(base)
    return z[prefix + "chunk_off"], ca, co


TABLES = os.path.join(GOLDEN_DIR, "tables_proc1200.npz")


def load_tables():
    z = np.load(TABLES)
    arena = sc.Arena(z["arena"], int(z["tlas_offset"]), [(int(o), int(s)) for o, s in z["blas"]])
    rays = z["rays"].view(_abi.RAY) if z["rays"].dtype != _abi.RAY else z["rays"]
    return z, arena, rays


COALESCING = os.path.join(GOLDEN_DIR, "coalescing_proc1500.npz")


def load_coalescing():
    """(fixture, arena, rays, expected transactions with BVH addresses rebased to the arena, expected stores)."""
    z = np.load(COALESCING)
    arena = sc.Arena(z["arena"], int(z["tlas_offset"]), [(int(o), int(s)) for o, s in z["blas"]])
    rays = z["rays"].view(_abi.RAY) if z["rays"].dtype != _abi.RAY else z["rays"]
    tx = np.zeros(len(z["txn_addr"]), _abi.TXN)
    tab = (z["txn_addr"] >> np.uint64(63)) != 0
    tx["address"] = np.where(tab, z["txn_addr"], z["txn_addr"] + np.uint64(arena.base)); tx["size"] = z["txn_size"]; tx["type"] = z["txn_type"]
    st = np.zeros(len(z["store_addr"]), _abi.STORE)
    st["address"] = z["store_addr"]; st["size"] = z["store_size"]
    return z, arena, rays, tx, st


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

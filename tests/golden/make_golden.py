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


if __name__ == "__main__":
    main()

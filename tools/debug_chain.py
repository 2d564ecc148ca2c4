import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers, vsrt.api as api
import test_gpu_baseline_scale as t
a = t.chain_arena(120)
for se in (96, 192):
    for mode in (0, 1):
        ctx = api.Context(max_treelet_size=512, device=0, stack_entries=se)
        ctx.register(a); ti = ctx.form_treelets()
        try:
            g = ctx.trace(mode, helpers.kat_ray(1))
            print("stack", se, "mode", mode, "records", len(g["txns"]), "hit", g["hits"]["hit_geometry"], g["hits"]["primitive_index"], "treelets", ti.n_treelets)
        except api.VsrtError as e:
            print("stack", se, "mode", mode, "error", e)
        ctx.close()

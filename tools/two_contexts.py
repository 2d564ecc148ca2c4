#!/usr/bin/env python
"""Does overlapping the stages of consecutive frames pay?  Two contexts on one GPU trace the bench frame concurrently from two host
threads (own streams): K1 of one frame can then share the SMs with scan + K3 of the other, and the host turn-around of one call is
covered by the other's kernels.  Prints frames/s for one context alone and for the pair.  Measurement only (profiles/README.md)."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "treelet-prefetching-for-rt_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from vsrt import scene as sc, _abi  # noqa: E402


def main():
    g.build()
    import vsrt.api as api
    dev = torch.device("cuda", 0)
    steps = int(os.environ.get("STEPS", "24"))
    s = sc.Scene(1_000_000, seed=0x5EED0001 + 1)
    rays = sc.rays_primary(1920, 1080, flags=0, tile=(8, 4))
    rd = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    ctxs = []
    for _ in range(2):
        c = api.Context(max_treelet_size=512, device=0)
        c.register(s); c.form_treelets()
        for _ in range(3):
            c.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))
        ctxs.append(c)
    torch.cuda.synchronize()

    def run(c, k):
        for _ in range(k):
            c.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))

    out = {}
    for name, group in (("one", ctxs[:1]), ("two", ctxs), ("one_again", ctxs[1:])):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(c, steps)) for c in group]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r = [c.device_results() for c in group]
        out[name] = {"frames_per_s": len(group) * steps / dt, "ms_per_frame": dt * 1e3 / (len(group) * steps),
                     "k1_ms_last": [x.traverse_ms for x in r], "k3_ms_last": [x.compact_ms for x in r]}
    # the persistent kernel's ramp and tail: K1 over two frames in ONE batch against twice K1 over one frame
    rd2 = torch.cat([rd, rd])
    c = ctxs[0]
    for _ in range(3):
        c.trace_device(_abi.MODE_TREELET, rd2.data_ptr(), 2 * len(rays))
    r2 = c.device_results()
    c.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))
    c.trace_device(_abi.MODE_TREELET, rd.data_ptr(), len(rays))
    r1 = c.device_results()
    out["two_frames_one_batch"] = {"k1_ms": r2.traverse_ms, "k3_ms": r2.compact_ms, "one_frame_k1_ms": r1.traverse_ms, "one_frame_k3_ms": r1.compact_ms}
    out["pair_over_single"] = out["two"]["frames_per_s"] / max(out["one"]["frames_per_s"], out["one_again"]["frames_per_s"])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

#!/bin/bash
# round 2, session 3: K1's traversal stack in global memory, [warp][entry][lane] x 16 B (VSRT_K1_GSTACK=1), against the local-memory stack
mkdir -p gpurun_out
L=treelet-prefetching-for-rt_b200/libvsrt_gst.so
VSRT_LIB=$L timeout 600 python -m pytest tests -m gpu -x -q -k "kat or random_scenes or stack_overflow or c2_bench or offset_quirk or node_layout or clustered" 2>&1 | tail -3
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M frac %.3f" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"]))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for V in "" _gst "" _gst; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
echo -n "bench DFS lib_gst: "; VSRT_BENCH_MODE=0 VSRT_LIB=$L $B 2>/dev/null | python -c "$J"
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for C in C3 C4; do for V in "" _gst; do echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"; done; done

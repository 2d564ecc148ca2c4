#!/bin/bash
# round 2, session 2: leaf cp.async (8 KB of shared memory per CTA -> 64 KB carve-out) against plain leaf loads (8 KB carve-out, 248 KB of L1)
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for V in "" _la0 "" _la0; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
for C in C3 C4; do for V in "" _la0; do
  echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done; done
for V in "" _la0; do echo -n "DFS lib$V: "; VSRT_BENCH_MODE=0 VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done

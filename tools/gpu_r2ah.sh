#!/bin/bash
# round 2, session 2: the layout-dispatch test; K1 loop knobs once more with the larger L1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "node_layout or scan_tile or packed_pipeline or kat" 2>&1 | tail -3
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for V in "" _i3 _t16 _t24 ""; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
for R in 6 12; do echo -n "bench refill=$R: "; VSRT_REFILL_T=$R $B 2>/dev/null | python -c "$J"; done
for L in 2 4; do echo -n "bench leaf=$L: "; VSRT_LEAF_T=$L $B 2>/dev/null | python -c "$J"; done
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for C in C3 C4; do for V in "" _t16; do echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"; done
  echo -n "$C refill=12: "; VSRT_REFILL_T=12 python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"; done

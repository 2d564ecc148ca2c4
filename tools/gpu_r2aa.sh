#!/bin/bash
# round 2, session 2: internal-node loads without L1 allocation, refill threshold on the incoherent configs
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for V in "" _nodena _tn0 _nodena_tn0; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
for C in C3 C4; do for V in "" _nodena _tn0 _nodena_tn0; do
  echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done; done
for C in C3 C4; do for R in 4 12 16 24; do
  echo -n "$C refill=$R: "; VSRT_REFILL_T=$R python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done; done

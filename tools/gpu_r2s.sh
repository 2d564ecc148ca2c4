#!/bin/bash
# round 2, session 2: where the ray-chunk K3 loses (histogram off / ILP / residency), L2 persistence window, allocation swap of the traversal copy
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
echo -n "v2: "; $B 2>/dev/null | python -c "$J"
echo -n "v2 nohist: "; VSRT_NO_HIST=1 $B 2>/dev/null | python -c "$J"
echo -n "windows nohist: "; VSRT_NO_HIST=1 VSRT_K3_WINDOWS=1 $B 2>/dev/null | python -c "$J"
echo -n "v2 ilp8: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_k3i8.so $B 2>/dev/null | python -c "$J"
echo -n "v2 blocks10: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_k3b10.so $B 2>/dev/null | python -c "$J"
for L in 16 48 96; do echo -n "windows l2persist=$L: "; VSRT_L2_PERSIST=$L VSRT_K3_WINDOWS=1 $B 2>/dev/null | python -c "$J"; done
echo -n "windows tnswap: "; VSRT_TN_SWAP=1 VSRT_K3_WINDOWS=1 $B 2>/dev/null | python -c "$J"
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f k3 %.3f |"%(p["k1_ms"],p.get("k3_ms",0)) for p in d["passes"]))'
for C in C3 C4; do
  echo -n "$C base: "; VSRT_K3_WINDOWS=1 python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
  echo -n "$C tnswap: "; VSRT_TN_SWAP=1 VSRT_K3_WINDOWS=1 python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
  echo -n "$C l2persist=64: "; VSRT_L2_PERSIST=64 VSRT_K3_WINDOWS=1 python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done

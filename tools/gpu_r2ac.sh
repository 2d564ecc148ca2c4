#!/bin/bash
# round 2, session 2: shared-memory carve-out of K1 (driver default = 132 KB for 64 KB used) -> L1 size
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for CV in -1 auto 43 14 -1 auto; do
  if [ $CV = auto ]; then unset VSRT_K1_CARVEOUT; else export VSRT_K1_CARVEOUT=$CV; fi
  echo -n "bench carveout=$CV: "; $B 2>/dev/null | python -c "$J"
done
for C in C3 C4; do for CV in -1 auto 43; do
  if [ $CV = auto ]; then unset VSRT_K1_CARVEOUT; else export VSRT_K1_CARVEOUT=$CV; fi
  echo -n "$C carveout=$CV: "; python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done; done
unset VSRT_K1_CARVEOUT
timeout 600 python -m pytest tests -m gpu -x -q -k "kat or random_scenes or golden or c2_bench" 2>&1 | tail -2
echo -n "DFS auto: "; VSRT_BENCH_MODE=0 $B 2>/dev/null | python -c "$J"
echo -n "DFS -1: "; VSRT_K1_CARVEOUT=-1 VSRT_BENCH_MODE=0 $B 2>/dev/null | python -c "$J"

#!/bin/bash
# round 2, session 3: order of the frame's ray ids -- scanlines vs the reference's raygen pixel tiles (WARP_8X4, vulkan_ray_tracing.cc:3505)
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M frac %.3f e2e %.1f M" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["e2e"]["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for T in 0 8x4 16x2 8x8 4x8; do echo -n "bench tile=$T: "; VSRT_BENCH_TILE=$T $B 2>/dev/null | python -c "$J"; done
for V in _t16 _t24 _i3; do echo -n "bench tile=8x4 lib$V: "; VSRT_BENCH_TILE=8x4 VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
for R in 12; do echo -n "bench tile=8x4 refill=$R: "; VSRT_BENCH_TILE=8x4 VSRT_REFILL_T=$R $B 2>/dev/null | python -c "$J"; done
echo -n "bench tile=8x4 DFS: "; VSRT_BENCH_TILE=8x4 VSRT_BENCH_MODE=0 $B 2>/dev/null | python -c "$J"

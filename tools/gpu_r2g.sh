#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_tests.log; cat gpurun_out/r2g_tests.log
for V in "" _leafasync; do for rep in 1 2; do
 echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M e2e_packed %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6, d['e2e_packed']['value']/1e6))"
done; done
echo -n "bench DFS lib: "; VSRT_BENCH_MODE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6))"
echo -n "bench DFS leafasync: "; VSRT_BENCH_MODE=0 VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_leafasync.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6))"
for C in C3 C4; do for B in 512 49152; do
  VSRT_K1_TB=1 python tools/prof_incoherent.py --config $C --budget $B --reps 2 2>&1 | tail -1 > gpurun_out/r2g_tb_${C}_$B.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2g_tb_${C}_$B.json")); print("TB $C $B:", [round(p["k1_ms"],2) for p in d["passes"]], d["tb_stats"])
PY
done; done

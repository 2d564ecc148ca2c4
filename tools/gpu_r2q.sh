#!/bin/bash
# round 2, session 2: traversal-layout internal nodes (VSRT_K1_TNODES) A/B + full GPU test suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2q_tests.log
for V in "" _tn0; do for rep in 1 2; do
 echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6))"
done; done
for C in C3 C4; do for V in "" _tn0; do
    echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
done; done
for V in "" _tn0; do echo -n "DFS lib$V: "; VSRT_BENCH_MODE=0 VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f' % b['k_traverse'])"; done

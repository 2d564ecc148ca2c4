#!/bin/bash
for V in "" _la2 _la2i3 _la2t16; do for rep in 1 2; do
 echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6))"
done; done
for C in C3 C4; do for V in "" _la2 _la2i3 _la2t16; do
    echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
done; done
echo -n "DFS lib: "; VSRT_BENCH_MODE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f' % b['k_traverse'])"
echo -n "DFS la2: "; VSRT_BENCH_MODE=0 VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_la2.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f' % b['k_traverse'])"

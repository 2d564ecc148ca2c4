#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned" 2>&1 | tail -15 > gpurun_out/r2e_tests.log; cat gpurun_out/r2e_tests.log
# knob sweeps in input order: C3 bounce and C4 bounce
for C in C3 C4; do for RT in 8 16 24; do for LT in 1 2 4; do
  echo -n "$C refill $RT leaf $LT: "; VSRT_RAY_ORDER=1 VSRT_REFILL_T=$RT VSRT_LEAF_T=$LT python tools/prof_incoherent.py --config $C --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
done; done; done
# the headline with the same knobs
for RT in 8 16; do for LT in 1 4; do
 echo -n "bench refill $RT leaf $LT: "; VSRT_REFILL_T=$RT VSRT_LEAF_T=$LT python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('k1 %.3f k3 %.3f value %.1f M' % (b['k_traverse'], b['k_compact'], d['value']/1e6))"
done; done
# treelet-binned variant on the incoherent configs
for B in 512 49152; do for TB in 0 1; do
  echo -n "C3 budget $B TB=$TB: "; VSRT_RAY_ORDER=1 VSRT_K1_TB=$TB python tools/prof_incoherent.py --config C3 --budget $B --reps 2 2>&1 | tail -1 | cut -c1-600
done; done

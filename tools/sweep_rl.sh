#!/bin/bash
# sweep refill / leaf thresholds of K1 (env knobs), print K1 ms
for r in ${REFILLS:-4 6 8 12}; do for l in ${LEAFS:-2 4 6 8 12}; do
VSRT_REFILL_T=$r VSRT_LEAF_T=$l python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']
print('refill',$r,'leaf',$l,'k1 %.3f ms'%b['k_traverse'])"
done; done

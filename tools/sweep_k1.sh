#!/bin/bash
# sweep K1 knobs on the GPU box: launch bounds (rebuild), refill / leaf thresholds (env)
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['roofline']['step_breakdown_ms']
print('$1 k1 %.3f ms k3 %.3f ms value %.1f Mrays/s'%(b['k_traverse'],b['k_compact'],d['value']/1e6))"; }
for mb in ${MINBLOCKS:-5 6 8}; do
  VSRT_NVCC_EXTRA="-DVSRT_K1_MIN_BLOCKS=$mb" python -c "import __graft_entry__ as g; g.build_cuda(force=True)"
  for r in ${REFILLS:-8}; do for l in ${LEAFS:-4}; do
    VSRT_REFILL_T=$r VSRT_LEAF_T=$l run "minblocks=$mb refill=$r leaf=$l"
  done; done
done

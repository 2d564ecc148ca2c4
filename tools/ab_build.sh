#!/bin/bash
# A/B on the GPU box: rebuild libvsrt.so with each set of extra nvcc flags and print the K1 time (2 bench runs each)
# usage: bash tools/ab_build.sh "-DX=0" "-DX=1" ...
for flags in "$@"; do
  VSRT_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" || exit 1
  for rep in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']
print('[$flags] k1 %.3f ms k3 %.3f ms value %.1f Mrays/s'%(b['k_traverse'],b['k_compact'],d['value']/1e6))"
  done
done

#!/bin/bash
# usage: VAR=NAME VALUES="a b c" bash tools/sweep_env.sh   -- bench once per value of one env knob
for v in $VALUES; do export $VAR=$v; python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']
print('$VAR=$v k1 %.3f ms k3 %.3f ms value %.1f Mrays/s'%(b['k_traverse'],b['k_compact'],d['value']/1e6))"; done

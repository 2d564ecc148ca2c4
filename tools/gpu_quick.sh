#!/bin/bash
# quick GPU check: parity tests, then the bench breakdown (K1 / K3 ms)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']
print('k1 %.3f ms k3 %.3f ms value %.1f Mrays/s frac %.3f'%(b['k_traverse'],b['k_compact'],d['value']/1e6,d['roofline']['frac']))"

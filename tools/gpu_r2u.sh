#!/bin/bash
# round 2, session 2: host-side expansion of the full form (VSRT_HOST_EXPAND, default on) -- tests, e2e A/B, worker-thread sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2u_tests.log
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; e=d["e2e"]; print("value %.1f M  e2e %.1f M (ok %s, d2h %.2f GB)  e2e_packed %.1f M" % (d["value"]/1e6, e["value"]/1e6, e.get("matches_device_records"), e["d2h_bytes_per_step"]/1e9, d["e2e_packed"]["value"]/1e6))'
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 3"
echo -n "expand=1: "; $B 2>gpurun_out/r2u_err.log | python -c "$J"; tail -3 gpurun_out/r2u_err.log
echo -n "expand=0: "; VSRT_HOST_EXPAND=0 $B 2>/dev/null | python -c "$J"
for T in 4 8 12 16 24 32; do echo -n "expand=1 threads=$T: "; VSRT_HOST_THREADS=$T $B 2>/dev/null | python -c "$J"; done
for C in 131072 262144 1048576; do echo -n "expand=1 chunk=$C: "; VSRT_PIPELINE_CHUNK=$C $B 2>/dev/null | python -c "$J"; done
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"; free -g | head -2

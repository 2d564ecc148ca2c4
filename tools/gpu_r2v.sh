#!/bin/bash
# round 2, session 2: worker pool + asynchronous window jobs in the host-buffer frame pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "packed or pipeline or empty_and_ragged or warp_call or kat" > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2v_tests.log
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; e=d["e2e"]; print("value %.1f M  e2e %.1f M (ok %s, d2h %.2f GB)  e2e_packed %.1f M" % (d["value"]/1e6, e["value"]/1e6, e.get("matches_device_records"), e["d2h_bytes_per_step"]/1e9, d["e2e_packed"]["value"]/1e6))'
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 4"
echo -n "expand=1: "; $B 2>gpurun_out/r2v_err.log | python -c "$J"; tail -3 gpurun_out/r2v_err.log
echo -n "expand=1: "; $B 2>/dev/null | python -c "$J"
echo -n "expand=0: "; VSRT_HOST_EXPAND=0 $B 2>/dev/null | python -c "$J"
for T in 12 15 20; do echo -n "expand=1 threads=$T: "; VSRT_HOST_THREADS=$T $B 2>/dev/null | python -c "$J"; done
for C in 196608 262144 393216; do echo -n "expand=1 chunk=$C: "; VSRT_PIPELINE_CHUNK=$C $B 2>/dev/null | python -c "$J"; done

#!/bin/bash
# round 2, session 2: node layout picked per batch on the device (both hot instantiations queued, gated on a coherence sample)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ab_tests.log
for L in 0 1; do VSRT_K1_LAYOUT=$L timeout 900 python -m pytest tests -m gpu -x -q -k "kat or random_scenes or c2_bench or c3_incoherent or golden" > gpurun_out/r2ab_tests_layout$L.log 2>&1; echo "tests(layout=$L) rc=$?"; tail -2 gpurun_out/r2ab_tests_layout$L.log; done
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M launches %d" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6, d["gpu_launches"]))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for L in auto 1 0 auto; do echo -n "bench layout=$L: "; if [ $L = auto ]; then $B 2>/dev/null | python -c "$J"; else VSRT_K1_LAYOUT=$L $B 2>/dev/null | python -c "$J"; fi; done
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f |"%(p["k1_ms"]) for p in d["passes"]))'
for C in C3 C4; do for L in auto 1 0; do
  echo -n "$C layout=$L: "; if [ $L = auto ]; then python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"; else VSRT_K1_LAYOUT=$L python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"; fi
done; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ab_bench.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M step %.3f k1 %.3f k3 %.3f frac %.3f | e2e %.1f M ok=%s | packed %.1f M | parity %s" % (d["value"]/1e6,d["ms_per_step"],b["k_traverse"],b["k_compact"],d["roofline"]["frac"],d["e2e"]["value"]/1e6,d["e2e"].get("matches_device_records"),d["e2e_packed"]["value"]/1e6, d.get("parity_sample",{}).get("equal")))
print("C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
for pb in i["C3"]["per_bounce"]: print("  C3 bounce", pb["bounce"], "k1 %.3f"%pb["k1_ms"])
PY

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_tests.log; cat gpurun_out/r2f_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 2000 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json")); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M k1 %.3f k3 %.3f frac %.3f | e2e %.1f M | e2e_packed %.1f M ok=%s | pcie %s" % (d["value"]/1e6,b["k_traverse"],b["k_compact"],d["roofline"]["frac"],d["e2e"]["value"]/1e6,d["e2e_packed"]["value"]/1e6,d["e2e_packed"].get("matches_device_records"),d.get("pcie")))
print("parity", d.get("parity_sample"))
print("C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
print(" C3 per bounce:", [(p["rays"], round(p["k1_ms"],3), round(p["rays_per_s"]/1e6,1)) for p in i["C3"]["per_bounce"]])
PY
for C in C3 C4; do
  echo -n "$C WF=1: "; VSRT_K1_WF=1 python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
  for V in "" _leafasync _pfleaf; do
    echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
  done
done
VSRT_K1_TB=1 python tools/prof_incoherent.py --config C3 --budget 49152 --reps 1 2>&1 | tail -1 > gpurun_out/r2f_tb_c3_48k.json
VSRT_K1_TB=1 python tools/prof_incoherent.py --config C4 --budget 49152 --reps 1 2>&1 | tail -1 > gpurun_out/r2f_tb_c4_48k.json
python tools/prof_incoherent.py --config C4 --budget 49152 --reps 2 2>&1 | tail -1 > gpurun_out/r2f_k1_c4_48k.json
cat gpurun_out/r2f_tb_c3_48k.json gpurun_out/r2f_tb_c4_48k.json gpurun_out/r2f_k1_c4_48k.json | cut -c1-1500

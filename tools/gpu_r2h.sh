#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2h_tests.log; cat gpurun_out/r2h_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench.json")); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M k1 %.3f k3 %.3f frac %.3f | e2e %.1f M | e2e_packed %.1f M ok=%s | pcie %.1f GB/s floors %.1f / %.1f ms" % (d["value"]/1e6,b["k_traverse"],b["k_compact"],d["roofline"]["frac"],d["e2e"]["value"]/1e6,d["e2e_packed"]["value"]/1e6,d["e2e_packed"].get("matches_device_records"),d["pcie"]["d2h_gbs_measured"],d["pcie"]["e2e_floor_ms"],d["pcie"]["e2e_packed_floor_ms"]))
print("parity", d.get("parity_sample",{}).get("equal"))
print("C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
print(" C3 per bounce:", [(p["rays"], round(p["k1_ms"],3), round(p["rays_per_s"]/1e6,1)) for p in i["C3"]["per_bounce"]])
PY
for T in 4 8 16 32; do echo -n "host threads $T: "; VSRT_HOST_THREADS=$T python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e %.1f M  e2e_packed %.1f M' % (d['e2e']['value']/1e6, d['e2e_packed']['value']/1e6))"; done
echo -n "no pipeline: "; VSRT_PIPELINE_CHUNK=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e %.1f M  e2e_packed %.1f M' % (d['e2e']['value']/1e6, d['e2e_packed']['value']/1e6))"
nproc; grep -m1 "model name" /proc/cpuinfo

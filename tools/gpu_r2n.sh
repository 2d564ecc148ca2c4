#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2n_tests.log; cat gpurun_out/r2n_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 800 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2n_bench.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M k1 %.3f k3 %.3f frac %.3f form %.2f ms | e2e %.1f M | e2e_packed %.1f M ok=%s" % (d["value"]/1e6,b["k_traverse"],b["k_compact"],d["roofline"]["frac"],d["config"]["treelet_form_ms"],d["e2e"]["value"]/1e6,d["e2e_packed"]["value"]/1e6,d["e2e_packed"].get("matches_device_records")))
print("parity", d.get("parity_sample",{}).get("equal"), "cpu", d.get("cpu_baseline",{}).get("value"))
print("C3 %.1f M frac %.3f form %.1f | C4 %.1f M frac %.3f k1 %.3f form %.1f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C3"]["treelet_form_ms"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"],i["C4"]["treelet_form_ms"]))
PY
for B in 512 49152; do python tools/prof_incoherent.py --config C4 --budget $B --reps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 form budget', d['budget'], 'form_ms', d['form_ms'], 'treelets', d['treelets'])"; done

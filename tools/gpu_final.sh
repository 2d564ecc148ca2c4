#!/bin/bash
# what the driver runs at the end of a round, on one box: GPU suite, smoke(), own arm with default flags
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/final_tests.log; [ $rc -ne 0 ] && exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; tail -c 300 gpurun_out/final_bench_1gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/final_bench_1gpu.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M step %.3f k1 %.3f scan %.4f k3 %.3f frac %.3f form %.2f ms launches %d | e2e %.1f M ok=%s | e2e_packed %.1f M ok=%s" % (d["value"]/1e6,d["ms_per_step"],b["k_traverse"],b["scan"],b["k_compact"],d["roofline"]["frac"],d["config"]["treelet_form_ms"],d["gpu_launches"],d["e2e"]["value"]/1e6,d["e2e"].get("matches_device_records"),d["e2e_packed"]["value"]/1e6,d["e2e_packed"].get("matches_device_records")))
print("parity", d.get("parity_sample",{}).get("equal"), "cpu", d.get("cpu_baseline",{}).get("value"), "clocks", d.get("clocks"))
print("C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
PY

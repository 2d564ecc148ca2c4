#!/bin/bash
mkdir -p gpurun_out
python tools/debug_chain.py 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c_tests.log; cat gpurun_out/r2c_tests.log
python tools/run_random.py | tee gpurun_out/r2c_random.jsonl
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("%s headline %.1f M k1 %.3f k3 %.3f | C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (sys.argv[2], d["value"]/1e6,b["k_traverse"],b["k_compact"],i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
print("   C3 per bounce:", [(p["rays"], round(p["k1_ms"],3), round(p["rays_per_s"]/1e6,1)) for p in i["C3"]["per_bounce"]])
PY
}
for V in "" _smem8 _smem12 _smem16; do
  VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2c_bench$V.json 2> gpurun_out/r2c_bench$V.err
  summ gpurun_out/r2c_bench$V.json "lib$V"
done
K=regex:k_traverseILi1ELi96ELb0
for C in C3 C4; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -c 3 -o gpurun_out/r2c_k1_$C -f python tools/prof_incoherent.py --config $C > gpurun_out/r2c_ncu_$C.log 2>&1
  bash tools/ncu_raw.sh gpurun_out/r2c_k1_$C.ncu-rep gpurun_out/r2c_k1_$C.raw.csv
  grep -E "^# kernel|gpu__time_duration.sum" gpurun_out/r2c_k1_$C.raw.csv
done

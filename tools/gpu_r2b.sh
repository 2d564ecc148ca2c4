#!/bin/bash
# round 2, second GPU call: GPU tests, ray-order A/B on the incoherent configs, K1 on the bounce batches under ncu, stack depths
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_tests.log; cat gpurun_out/r2b_tests.log
for RO in 1 0; do
  VSRT_RAY_ORDER=$RO python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2b_bench_order$RO.json 2> gpurun_out/r2b_bench_order$RO.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b_bench_order$RO.json")); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("ray_order=$RO headline %.1f M k1 %.3f k3 %.3f | C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (d["value"]/1e6,b["k_traverse"],b["k_compact"],i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
print(" C3 per bounce:", [(p["rays"], round(p["k1_ms"],3), round(p["rays_per_s"]/1e6,1)) for p in i["C3"]["per_bounce"]])
PY
done
K=regex:k_traverseILi1ELi96ELb0
for C in C3 C4; do
  VSRT_RAY_ORDER=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 1 -c 1 -o gpurun_out/r2b_k1_$C -f python tools/prof_incoherent.py --config $C > gpurun_out/r2b_ncu_$C.log 2>&1
  tail -1 gpurun_out/r2b_ncu_$C.log | cut -c1-400
  bash tools/ncu_raw.sh gpurun_out/r2b_k1_$C.ncu-rep gpurun_out/r2b_k1_$C.raw.csv
done
VSRT_RAY_ORDER=2 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 1 -c 1 -o gpurun_out/r2b_k1_C4_sorted -f python tools/prof_incoherent.py --config C4 > gpurun_out/r2b_ncu_C4_sorted.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2b_k1_C4_sorted.ncu-rep gpurun_out/r2b_k1_C4_sorted.raw.csv
VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_stats.so python tools/k1_depth.py bench 512 | tee gpurun_out/r2b_depth_bench.json
VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_stats.so python tools/k1_depth.py c3 512 | tee gpurun_out/r2b_depth_c3.json
VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_stats.so python tools/k1_depth.py bench 49152 | tee gpurun_out/r2b_depth_bench48k.json

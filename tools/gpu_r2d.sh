#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2d_tests.log; cat gpurun_out/r2d_tests.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("%s headline %.1f M k1 %.3f k3 %.3f | C3 %.1f M frac %.3f | C4 %.1f M frac %.3f k1 %.3f" % (sys.argv[2], d["value"]/1e6,b["k_traverse"],b["k_compact"],i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"]))
print("   C3 per bounce:", [(p["rays"], round(p["k1_ms"],3), round(p["rays_per_s"]/1e6,1)) for p in i["C3"]["per_bounce"]])
PY
}
for V in "" _pfl2 _pfl1 _mb8; do
  VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2d_bench$V.json 2> gpurun_out/r2d_bench$V.err
  summ gpurun_out/r2d_bench$V.json "lib$V"
done
# knob sweep on the C4 bounce batch (no ncu): refill x leaf thresholds
for RT in 4 8 16; do for LT in 1 2 4 8; do
  echo -n "refill $RT leaf $LT: "; VSRT_REFILL_T=$RT VSRT_LEAF_T=$LT python tools/prof_incoherent.py --config C4 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(' '.join('%.3f'%p['k1_ms'] for p in d['passes']))"
done; done
# why does sorting not pay?  K1 on 2 M random rays, input order vs sorted, under ncu (first hot launch of each context)
K=regex:k_traverseILi1ELi96ELb0
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -c 9 -o gpurun_out/r2d_k1_random -f python tools/run_random.py > gpurun_out/r2d_ncu_random.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2d_k1_random.ncu-rep gpurun_out/r2d_k1_random.raw.csv
grep -E "^# kernel|gpu__time_duration.sum" gpurun_out/r2d_k1_random.raw.csv

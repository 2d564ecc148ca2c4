#!/bin/bash
# round 2, session 2: why the traversal copy (VSRT_K1_TNODES) loses on the incoherent configs -- K1 on C4 under ncu, both builds; K3 HIST2 variant
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
timeout 600 python -m pytest tests -m gpu -x -q -k "kat or random_scenes or golden or histogram or packed or c2_bench" > gpurun_out/r2t_tests.log 2>&1; echo "tests(default) rc=$?"; tail -2 gpurun_out/r2t_tests.log
VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_hist2.so timeout 600 python -m pytest tests -m gpu -x -q -k "kat or random_scenes or golden or histogram or packed or c2_bench" > gpurun_out/r2t_tests_hist2.log 2>&1; echo "tests(hist2) rc=$?"; tail -2 gpurun_out/r2t_tests_hist2.log
for V in "" _hist2 "" _hist2; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
K=regex:k_traverseILi1ELi96ELb0
for V in "" _tn0; do
  VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -c 4 -o gpurun_out/r2t_k1_C4$V -f python tools/prof_incoherent.py --config C4 > gpurun_out/r2t_ncu_C4$V.log 2>&1
  bash tools/ncu_raw.sh gpurun_out/r2t_k1_C4$V.ncu-rep gpurun_out/r2t_k1_C4$V.raw.csv
  ncu -i gpurun_out/r2t_k1_C4$V.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2t_k1_C4${V}_src.csv 2>/dev/null
done
for f in r2t_k1_C4 r2t_k1_C4_tn0; do echo "== $f"; grep -E "^# kernel|gpu__time_duration.sum,|dram__bytes_read.sum,|dram__bytes_write.sum,|l1tex__t_sector_hit|lts__t_sector_hit|smsp__inst_executed.sum,|thread_inst_executed_per_inst|issue_active.avg.pct_of_peak_sustained_active|long_scoreboard" gpurun_out/$f.raw.csv; done
# L1 cache hints for K1 (leaf copies bypass L1, internal-node loads evict-last) and refill threshold with the traversal copy
for V in "" _leafcg _nodeel _cgel; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
for R in 6 10 12; do echo -n "bench refill=$R: "; VSRT_REFILL_T=$R $B 2>/dev/null | python -c "$J"; done
P='import json,sys
d=json.loads(sys.stdin.read()); print(" ".join("k1 %.3f k3 %.3f |"%(p["k1_ms"],p.get("k3_ms",0)) for p in d["passes"]))'
for C in C3 C4; do for V in "" _leafcg _nodeel _cgel; do
  echo -n "$C lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so python tools/prof_incoherent.py --config $C --reps 3 2>&1 | tail -1 | python -c "$P"
done; done

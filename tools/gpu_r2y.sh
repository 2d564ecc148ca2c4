#!/bin/bash
# end-of-round captures at the committed state (second pass, after the layout dispatch and the carve-out): tests, own arm, launch list, K1 / K3 under ncu --set full
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y2_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2y2_tests.log; [ $rc -ne 0 ] && exit 1
python bench.py > gpurun_out/r2y2_bench_1gpu.json 2> gpurun_out/r2y2_bench_1gpu.err; tail -c 400 gpurun_out/r2y2_bench_1gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2y2_bench_1gpu.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]; i=d.get("incoherent",{})
print("headline %.1f M step %.3f k1 %.3f scan %.4f k3 %.3f frac %.3f form %.2f ms | e2e %.1f M ok=%s | e2e_packed %.1f M ok=%s" % (d["value"]/1e6,d["ms_per_step"],b["k_traverse"],b["scan"],b["k_compact"],d["roofline"]["frac"],d["config"]["treelet_form_ms"],d["e2e"]["value"]/1e6,d["e2e"].get("matches_device_records"),d["e2e_packed"]["value"]/1e6,d["e2e_packed"].get("matches_device_records")))
print("parity", d.get("parity_sample",{}).get("equal"), "cpu", d.get("cpu_baseline",{}).get("value"), "clocks", d.get("clocks"))
print("C3 %.1f M frac %.3f form %.1f | C4 %.1f M frac %.3f k1 %.3f form %.1f" % (i["C3"]["value"]/1e6,i["C3"]["roofline"]["frac"],i["C3"]["treelet_form_ms"],i["C4"]["value"]/1e6,i["C4"]["roofline"]["frac"],i["C4"]["k1_ms"],i["C4"]["treelet_form_ms"]))
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y2_launches.csv $B > gpurun_out/r2y2_launches.log 2>&1
K=regex:k_traverseILi1ELi96ELb0ELb1     # the hot instantiation over the traversal copy (the Mesa-layout one is queued too and returns at once on this workload)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 4 -c 1 -o gpurun_out/r2y2_k1_bench -f $B > gpurun_out/r2y2_ncu_k1.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2y2_k1_bench.ncu-rep gpurun_out/r2y2_k1_bench.raw.csv
ncu --set full --clock-control none --import-source on -k regex:k_compact -s 4 -c 1 -o gpurun_out/r2y2_k3_bench -f $B > gpurun_out/r2y2_ncu_k3.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2y2_k3_bench.ncu-rep gpurun_out/r2y2_k3_bench.raw.csv
for f in r2y2_k1_bench r2y2_k3_bench; do echo "== $f"; grep -E "^# kernel|gpu__time_duration.sum,|dram__bytes_read.sum,|dram__bytes_write.sum,|l1tex__t_sector_hit|lts__t_sector_hit|smsp__inst_executed.sum,|thread_inst_executed_per_inst|issue_active.avg.pct_of_peak_sustained_active|long_scoreboard|pipe_alu" gpurun_out/$f.raw.csv; done
timeout 1500 python tools/run_configs.py --configs C3,C4,C5 > gpurun_out/r2y2_configs_C3_C4_C5.jsonl 2> gpurun_out/r2y2_configs.err; tail -c 300 gpurun_out/r2y2_configs.err; cut -c1-200 gpurun_out/r2y2_configs_C3_C4_C5.jsonl
VSRT_BENCH_MODE=0 $B 2>/dev/null > gpurun_out/r2y2_bench_dfs_mode.json; python -c "
import json; d=json.load(open('gpurun_out/r2y2_bench_dfs_mode.json')); b=d['roofline']['step_breakdown_ms']; print('DFS mode: value %.1f M k1 %.3f k3 %.3f frac %.3f'%(d['value']/1e6,b['k_traverse'],b['k_compact'],d['roofline']['frac']))"

#!/bin/bash
# round 2, session 3: K1 / K3 under ncu --set full at the committed state (8x4 ray order, second round from 16 lanes), reference arm, traceRay mode
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
K=regex:k_traverseILi1ELi96ELb0ELb1     # the hot instantiation over the traversal copy (the Mesa-layout one is queued too and returns at once on this workload)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 4 -c 1 -o gpurun_out/r2c_k1_bench -f $B > gpurun_out/r2c_ncu_k1.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2c_k1_bench.ncu-rep gpurun_out/r2c_k1_bench.raw.csv
ncu --set full --clock-control none --import-source on -k regex:k_compact -s 4 -c 1 -o gpurun_out/r2c_k3_bench -f $B > gpurun_out/r2c_ncu_k3.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2c_k3_bench.ncu-rep gpurun_out/r2c_k3_bench.raw.csv
for f in r2c_k1_bench r2c_k3_bench; do echo "== $f"; grep -E "^# kernel|gpu__time_duration.sum,|dram__bytes_read.sum,|dram__bytes_write.sum,|l1tex__t_sector_hit|lts__t_sector_hit|smsp__inst_executed.sum,|thread_inst_executed_per_inst|issue_active.avg.pct_of_peak_sustained_active|long_scoreboard|pipe_alu" gpurun_out/$f.raw.csv; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_ref.err; tail -c 700 gpurun_out/r2c_bench_reference_arm.json
VSRT_BENCH_MODE=0 $B 2>/dev/null > gpurun_out/r2c_bench_dfs_mode.json; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_dfs_mode.json')); b=d['roofline']['step_breakdown_ms']; print('DFS mode: value %.1f M k1 %.3f k3 %.3f frac %.3f'%(d['value']/1e6,b['k_traverse'],b['k_compact'],d['roofline']['frac']))"

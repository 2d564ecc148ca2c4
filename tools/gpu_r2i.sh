#!/bin/bash
# final captures of the round: launch list of the bench step, K1 / K3 under ncu --set full (headline), K1 on the C3 bounce batch in
# input order, the treelet-binned round kernel; raw-page extracts for profiles/
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv $B > gpurun_out/r2i_launches.log 2>&1
K=regex:k_traverseILi1ELi96ELb0
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 4 -c 1 -o gpurun_out/r2i_k1_bench -f $B > gpurun_out/r2i_ncu_k1.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2i_k1_bench.ncu-rep gpurun_out/r2i_k1_bench.raw.csv
ncu --set full --clock-control none --import-source on -k regex:k_compact -s 4 -c 1 -o gpurun_out/r2i_k3_bench -f $B > gpurun_out/r2i_ncu_k3.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2i_k3_bench.ncu-rep gpurun_out/r2i_k3_bench.raw.csv
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -c 4 -o gpurun_out/r2i_k1_C3 -f python tools/prof_incoherent.py --config C3 > gpurun_out/r2i_ncu_C3.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2i_k1_C3.ncu-rep gpurun_out/r2i_k1_C3.raw.csv
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -c 4 -o gpurun_out/r2i_k1_C4 -f python tools/prof_incoherent.py --config C4 > gpurun_out/r2i_ncu_C4.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2i_k1_C4.ncu-rep gpurun_out/r2i_k1_C4.raw.csv
VSRT_K1_TB=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tb_round -s 1 -c 2 -o gpurun_out/r2i_tb_round -f python tools/prof_incoherent.py --config C3 --budget 49152 --reps 1 > gpurun_out/r2i_ncu_tb.log 2>&1
bash tools/ncu_raw.sh gpurun_out/r2i_tb_round.ncu-rep gpurun_out/r2i_tb_round.raw.csv
for f in r2i_k1_bench r2i_k3_bench r2i_k1_C3 r2i_k1_C4 r2i_tb_round; do echo "== $f"; grep -E "^# kernel|gpu__time_duration.sum,|dram__bytes_read.sum,|dram__bytes_write.sum,|l1tex__t_sector_hit|lts__t_sector_hit|thread_inst_executed_per_inst|issue_active.avg.pct_of_peak_sustained_active|long_scoreboard|wavefronts_mem_shared.sum," gpurun_out/$f.raw.csv; done

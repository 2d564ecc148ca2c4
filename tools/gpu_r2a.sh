#!/bin/bash
# round 2, first GPU call: full GPU test suite, the bench line, launch list, K1 on the incoherent configs under ncu --set full
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.err
K=regex:k_traverseILi1ELi96ELb0
for C in C3 C4; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $K -s 1 -c 1 -o gpurun_out/r2a_k1_$C -f python tools/prof_incoherent.py --config $C > gpurun_out/r2a_ncu_$C.log 2>&1
  tail -1 gpurun_out/r2a_ncu_$C.log | cut -c1-400
  bash tools/ncu_raw.sh gpurun_out/r2a_k1_$C.ncu-rep gpurun_out/r2a_k1_$C.raw.csv
done
cat gpurun_out/r2a_tests.log
cat gpurun_out/r2a_bench.json | cut -c1-6000

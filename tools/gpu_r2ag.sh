#!/bin/bash
# round 2, session 2: does the reduce's NCCL kernel cost K1 its L1 carve-out at N > 2?  bench at N ranks, default vs few NCCL CTAs
mkdir -p gpurun_out
N=${1:-4}
for CT in default 4; do
  if [ $CT = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$CT; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 5 --e2e-steps 1 > gpurun_out/r2ag_bench_${N}gpu_$CT.json 2> gpurun_out/r2ag_${N}gpu_$CT.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2ag_bench_${N}gpu_$CT.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]
print("N=%d NCCL_MAX_CTAS=$CT value %.1f M ms/step %.3f k1 %.3f k3 %.3f reduce_ok %s" % (d["n_gpus"], d["value"]/1e6, d["ms_per_step"], b["k_traverse"], b["k_compact"], d["reduced_counters"]["reduce_ok"]))
PY
done

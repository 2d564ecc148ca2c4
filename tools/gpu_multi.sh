#!/bin/bash
# N GPUs of one box: the library-side reduce test, then the bench at N (weak scaling, totals asserted)
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_gpu_baseline_scale.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 5 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err
tail -c 800 gpurun_out/r2c_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2c_bench_${N}gpu.json") if l.startswith("{")][-1]); b=d["roofline"]["step_breakdown_ms"]
print("N=%d value %.1f M ms/step %.3f k1 %.3f k3 %.3f e2e %.1f M (ok %s) e2e_packed %.1f M" % (d["n_gpus"], d["value"]/1e6, d["ms_per_step"], b["k_traverse"], b["k_compact"], d["e2e"]["value"]/1e6, d["e2e"].get("matches_device_records"), d["e2e_packed"]["value"]/1e6))
print(d.get("reduced_counters"))
PY

#!/bin/bash
# the driver's round-end sequence on one box: gpu tests, smoke, reference arm, own arm (default flags)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_$1_ref.json 2> gpurun_out/bench_$1_ref.err; tail -c 1200 gpurun_out/bench_$1_ref.json
python bench.py > gpurun_out/bench_$1_1gpu.json 2> gpurun_out/bench_$1_1gpu.err; cat gpurun_out/bench_$1_1gpu.json

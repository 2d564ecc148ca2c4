#!/bin/bash
# round 2, session 2: loop knobs of K1 with the traversal copy; full bench line (incoherent object included) with and without it
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for V in "" _i3 _t16 _t24 ""; do echo -n "bench lib$V: "; VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt$V.so $B 2>/dev/null | python -c "$J"; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_tn.json 2> gpurun_out/r2w_bench_tn.err; tail -2 gpurun_out/r2w_bench_tn.err
VSRT_LIB=treelet-prefetching-for-rt_b200/libvsrt_tn0.so python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_tn0.json 2> gpurun_out/r2w_bench_tn0.err
for f in tn tn0; do python - $f <<'P'
import json,sys
d=json.load(open('gpurun_out/r2w_bench_%s.json'%sys.argv[1])); i=d['incoherent']
print(sys.argv[1], 'value %.1f M  e2e %.1f  packed %.1f  parity %s | C3 %.1f M (frac %.3f)  C4 %.1f M (frac %.3f)'%(d['value']/1e6, d['e2e']['value']/1e6, d['e2e_packed']['value']/1e6, d.get('parity_sample',{}).get('equal'), i['C3']['value']/1e6, i['C3']['roofline']['frac'], i['C4']['value']/1e6, i['C4']['roofline']['frac']))
P
done

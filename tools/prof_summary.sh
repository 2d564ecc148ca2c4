#!/bin/bash
# usage: tools/prof_summary.sh NAME [top_lines]   -- reads gpurun_out/prof_traverse_NAME.ncu-rep: opcode mix, hot lines, stall mix
cd "$(dirname "$0")/../gpurun_out" || exit 1
n=$1
ncu -i prof_traverse_$n.ncu-rep --page source --csv --print-source sass > ${n}_sass.csv 2>/dev/null
python ../tools/ncu_ops.py ${n}_sass.csv 2>/dev/null | head -${3:-14}
ncu -i prof_traverse_$n.ncu-rep --page source --csv --print-source cuda,sass > ${n}_src.csv 2>/dev/null
python ../tools/ncu_lines.py ${n}_src.csv ${2:-30}
ncu -i prof_traverse_$n.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
d=dict(zip(h,v))
for k in h:
    if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
        try:
            if float(d[k])>0.2: print('%-40s %s'%(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),d[k]))
        except: pass
print('ms', d['gpu__time_duration.sum'], 'lanes', d['smsp__thread_inst_executed_per_inst_executed.ratio'], 'alu%', d['sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'], 'issue%', d.get('sm__inst_issued.avg.pct_of_peak_sustained_active', d.get('smsp__issue_active.avg.pct_of_peak_sustained_active')), 'L1hit', d['l1tex__t_sector_hit_rate.pct'])
"

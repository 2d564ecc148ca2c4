#!/usr/bin/env python
"""Per-source-line summary of an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export:
warp-instructions executed, average active lanes, share of stall samples.  usage: ncu_lines.py export.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []; hdr = None; cur_file = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); ss = int(d["# Samples"])
    except (ValueError, KeyError):
        continue
    lsb = int(d.get("stall_long_sb") or 0)
    out.append((cur_file, int(r[0]), r[1].strip()[:100], ie, te, ss, lsb))
tot_i = sum(o[3] for o in out); tot_s = sum(o[5] for o in out)
print("total warp-instr %.3f G, thread-instr %.3f G (%.1f lanes), samples %d" % (tot_i / 1e9, sum(o[4] for o in out) / 1e9, sum(o[4] for o in out) / max(tot_i, 1), tot_s))
print("%-20s %8s %6s %6s %6s %6s  %s" % ("file:line", "Minstr", "%instr", "lanes", "%smpl", "%longsb", "source"))
key = (lambda o: -o[5]) if (len(sys.argv) > 3 and sys.argv[3] == "samples") else (lambda o: -o[3])
for o in sorted(out, key=key)[:top]:
    if o[3] == 0: continue
    print("%-20s %8.1f %6.2f %6.1f %6.2f %6.2f  %s" % ("%s:%d" % (o[0][:14], o[1]), o[3] / 1e6, 100.0 * o[3] / max(tot_i, 1), o[4] / max(o[3], 1), 100.0 * o[5] / max(tot_s, 1), 100.0 * o[6] / max(tot_s, 1), o[2]))

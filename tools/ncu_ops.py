#!/usr/bin/env python
"""Dynamic instruction mix of a kernel from an `ncu --page source --csv --print-source sass` export: executed warp
instructions per opcode and per issue pipe.  usage: ncu_ops.py export.csv"""
import csv, sys, re, collections
sys.path.insert(0, __import__("os").path.dirname(__file__))
ALU = {"LOP3","SHF","PRMT","SEL","ISETP","FSETP","FMNMX","FMNMX3","IADD3","LEA","MOV","VIADD","FSEL","PLOP3","IABS","SGXT","BMSK","VIMNMX","VIMNMX3","IMNMX","LOP","FSET","CS2R","P2R","R2P","FCHK","UMOV","R2UR","VOTE","MATCH","SHFL"}
FMA = {"FFMA","FMUL","FADD","IMAD","FADD2","FMUL2","FFMA2","HFMA2","FHADD","HADD2","HMUL2","FHFMA"}
XU = {"MUFU","POPC","FLO","I2F","F2I","BREV","I2FP","F2F"}
LSU = {"LDG","STG","LDL","STL","LDS","STS","ATOMS","ATOMG","REDG","RED","ATOM","LDC","LDCU","CCTL"}
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; ops = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] in ("Address", "Line No"): hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    src = d.get("Source", "")
    m = re.match(r"\s*(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", src)
    if not m: continue
    try: ie = int(d["Instructions Executed"])
    except (ValueError, KeyError): continue
    ops[m.group(1)] += ie
tot = sum(ops.values()); pipes = collections.Counter()
for o, n in ops.items():
    pipes["alu" if o in ALU else "fma" if o in FMA else "xu" if o in XU else "lsu" if o in LSU else "ctl/other"] += n
print("total %.3f G warp-instr" % (tot / 1e9))
for p, n in pipes.most_common(): print("  pipe %-10s %8.1f M  %5.1f %%" % (p, n / 1e6, 100.0 * n / tot))
for o, n in ops.most_common(40): print("  %-10s %8.1f M  %5.1f %%" % (o, n / 1e6, 100.0 * n / tot))

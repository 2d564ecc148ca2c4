#!/usr/bin/env python
"""Static count of SASS instructions per issue pipe between two addresses of a kernel dump (cuobjdump -sass).
usage: sass_pipes.py k1.sass [lo_hex hi_hex]"""
import re, sys, collections
ALU = {"LOP3","SHF","PRMT","SEL","ISETP","FSETP","FMNMX","FMNMX3","IADD3","LEA","MOV","VIADD","FSEL","PLOP3","IABS","SGXT","BMSK","VIMNMX","VIMNMX3","IMNMX","LOP","FSET","CS2R","P2R","R2P","VABSDIFF","FCHK","UMOV","R2UR"}
FMA = {"FFMA","FMUL","FADD","IMAD","FADD2","FMUL2","FFMA2","HFMA2","FHADD","HADD2","HMUL2","FHFMA"}
XU = {"MUFU","POPC","FLO","I2F","F2I","BREV","I2FP","F2F"}
LSU = {"LDG","STG","LDL","STL","LDS","STS","ATOMS","ATOMG","REDG","RED","ATOM","LDC","LDCU","CCTL"}
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
c = collections.Counter(); ops = collections.Counter()
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", line)
    if not m: continue
    a = int(m.group(1), 16)
    if a < lo or a > hi: continue
    op = m.group(2)
    pipe = "alu" if op in ALU else "fma" if op in FMA else "xu" if op in XU else "lsu" if op in LSU else "ctl"
    c[pipe] += 1; ops[(pipe, op)] += 1
print(dict(c), "total", sum(c.values()))
for (p, o), n in sorted(ops.items(), key=lambda kv: (kv[0][0], -kv[1])): print("  %-4s %-8s %d" % (p, o, n))

#!/usr/bin/env python
"""Summarise an ncu report's source page per CUDA line: usage  ncu_lines.py report.ncu-rep [topN] [kernel-name]"""
import csv, subprocess, sys
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 25
cmd=['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass']
if len(sys.argv)>3: cmd+=['--kernel-name',sys.argv[3]]
txt=subprocess.run(cmd,stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
h=None; lines=[]
for r in rows:
    if r and r[0]=='Line No' and len(r)>5:
        if '# Samples' not in r: h=None; continue
        h=r; S=h.index('# Samples'); I=h.index('Instructions Executed'); continue
    if h and len(r)==len(h) and r[0] not in ('','Line No'):
        try: lines.append((int(r[0]),r[1],int(r[S] or 0),int(r[I] or 0)))
        except ValueError: pass
tot=sum(l[2] for l in lines) or 1; toti=sum(l[3] for l in lines) or 1
print("samples=%d warp_instructions=%d"%(tot,toti))
for l in sorted(lines,key=lambda l:-l[2])[:top]:
    print(str(l[0]).rjust(4), ('%.1f%%'%(100*l[2]/tot)).rjust(6), ('%.1f%%'%(100*l[3]/toti)).rjust(6), l[1].strip()[:120])

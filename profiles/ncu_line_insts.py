import csv,subprocess,sys
rep=sys.argv[1]; top=int(sys.argv[2]); tiles=float(sys.argv[3])
txt=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
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
print("total inst", toti, "per tile", toti/tiles)
for l in sorted(lines,key=lambda l:-l[3])[:top]:
    print(str(l[0]).rjust(4), ('%5.1f%% inst'%(100*l[3]/toti)), ('%5.1f%% smp'%(100*l[2]/tot)), ('%6.1f/tile'%(l[3]/tiles)), l[1].strip()[:100])

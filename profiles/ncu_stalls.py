import csv,subprocess,sys
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
txt=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
h=rows[1]
S=h.index('# Samples'); I=h.index('Instructions Executed')
st=[c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
idx={c:h.index(c) for c in st}
tot={c:0 for c in st}; n=0; ninst=0
data=[]
for r in rows[2:]:
    if len(r)!=len(h): continue
    try: s=int(r[S] or 0)
    except: continue
    n+=s; ninst+=int(r[I] or 0)
    d={c:int(r[idx[c]] or 0) for c in st}
    for c in st: tot[c]+=d[c]
    data.append((s,r[0],r[1],d,int(r[I] or 0)))
print("samples",n,"inst",ninst)
for c,v in sorted(tot.items(), key=lambda x:-x[1])[:9]: print("%-28s %6.1f%%"%(c,100*v/n))
print()
for s,a,src,d,i in sorted(data,key=lambda x:-x[0])[:top]:
    t=sorted(d.items(), key=lambda x:-x[1])[:2]
    print("%5.2f%% %9d %-72s %s"%(100*s/n, i, src[:72], " ".join("%s=%d"%(k[6:],v) for k,v in t)))

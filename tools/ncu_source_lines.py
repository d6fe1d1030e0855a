import csv, sys, collections
rows = list(csv.reader(sys.stdin))
fn=None; fpath=None; hdr=None; last_line="0"
agg=collections.defaultdict(lambda: collections.defaultdict(lambda:[0,0,'']))
for r in rows:
    if len(r)>=2 and r[0]=='File Path': fpath=r[1]; continue
    if len(r)>=2 and r[0]=='Function Name': fn=r[1]; hdr=None; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr):
        # columns: Line No, Source(cuda), Address, Source(sass), ...
        line=r[0] or last_line; last_line=line
        try: ns=int(r[hdr.index('# Samples')]); ex=int(r[hdr.index('Instructions Executed')])
        except: continue
        a=agg[fn][(fpath.split('/')[-1],int(line))]
        a[0]+=ns; a[1]+=ex; a[2]=r[1]
for fn,d in agg.items():
    tot=sum(v[0] for v in d.values())
    print('=====',fn[:100],'samples',tot)
    for (f,l),v in sorted(d.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[1]) if len(sys.argv)>1 else 30]:
        print(f"{v[0]:7d} {100*v[0]/max(tot,1):5.1f}% ex={v[1]:9d} {f}:{l}: {v[2].strip()[:100]}")

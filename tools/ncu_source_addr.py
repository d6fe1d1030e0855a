import csv, sys, collections
rows = list(csv.reader(sys.stdin))
fn=None; fpath=None; hdr=None; last='0'
per=collections.defaultdict(list)
for r in rows:
    if len(r)>=2 and r[0]=='File Path': fpath=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=='Function Name': fn=r[1]; hdr=None; continue
    if r and r[0]=='Line No': hdr=r; idx={h:i for i,h in enumerate(hdr)}; continue
    if hdr and len(r)==len(hdr):
        line=r[0] or last; last=line
        try:
            per[fn].append((int(r[2],16), fpath, int(line), int(r[idx['# Samples']]), int(r[idx['Instructions Executed']]), r[3].strip(), int(r[idx['stall_long_sb']]), int(r[idx['stall_short_sb']]), int(r[idx['stall_mio']]),int(r[idx['stall_wait']])))
        except Exception as e: pass
want=sys.argv[1]
for fn,v in per.items():
    if want not in fn: continue
    v.sort()
    base=v[0][0]
    # run-length by role anchor: bucket of 64 instructions
    print('=====',fn[:100], len(v))
    B=int(sys.argv[2]) if len(sys.argv)>2 else 48
    for i in range(0,len(v),B):
        ch=v[i:i+B]
        n=sum(c[3] for c in ch)
        if n<int(sys.argv[3] if len(sys.argv)>3 else 150): continue
        lines=collections.Counter()
        for c in ch: lines[(c[1][:12],c[2])]+=c[3]
        top=' '.join(f"{f}:{l}={s}" for (f,l),s in lines.most_common(4))
        print(f"+{ch[0][0]-base:6x} n={n:6d} lsb={sum(c[6] for c in ch):5d} ssb={sum(c[7] for c in ch):5d} mio={sum(c[8] for c in ch):5d} wait={sum(c[9] for c in ch):5d} ex={sum(c[4] for c in ch):9d} | {top}")

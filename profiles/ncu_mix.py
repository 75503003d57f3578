import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; 
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
stall_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops=collections.Counter(); samp=collections.Counter(); tot=0; ts=0
stalls=collections.Counter()
for r in rows[2:]:
    if len(r)<=iE: continue
    src=r[iS].strip(); 
    toks=src.split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    op=op.split('.')[0]
    n=int(r[iE]); s=int(r[iSm])
    ops[op]+=n; samp[op]+=s; tot+=n; ts+=s
    for i in stall_cols: stalls[hdr[i]]+=int(r[i] or 0)
print("total inst",tot,"samples",ts)
for op,n in ops.most_common(30): print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[op]/ts:5.1f}%")
print(sorted(stalls.items(), key=lambda x:-x[1])[:8])

import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
sc={h[6:]:i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
B=int(sys.argv[2])
acc={}
for k,r in enumerate(rows[2:]):
    if len(r)<=iE: continue
    b=k//B
    a=acc.setdefault(b,dict(inst=0,samp=0,no_inst=0,long_sb=0,wait=0,short_sb=0,math=0,first=r[iS].strip()[:30]))
    a['inst']+=int(r[iE]); a['samp']+=int(r[iSm])
    for n in ('no_inst','long_sb','wait','short_sb','math'): a[n]+=int(r[sc[n]] or 0)
for b,a in acc.items(): print(b*B, a)

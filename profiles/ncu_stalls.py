import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; iSm=hdr.index("# Samples")
sc={h[6:]:i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
lo=int(sys.argv[2]) if len(sys.argv)>2 else 0; hi=int(sys.argv[3]) if len(sys.argv)>3 else 10**9
tot={k:0 for k in sc}; ts=0
for k,r in enumerate(rows[2:]):
    if k<lo or k>=hi or len(r)<=iSm: continue
    ts+=int(r[iSm])
    for n,i in sc.items(): tot[n]+=int(r[i] or 0)
print("samples",ts, {k:f"{100*v/ts:.1f}%" for k,v in sorted(tot.items(), key=lambda x:-x[1]) if v})

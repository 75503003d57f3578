import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
sc={h:i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
lo=int(sys.argv[2]); hi=int(sys.argv[3]); thr=int(sys.argv[4])
for k,r in enumerate(rows[2:]):
    if k<lo or k>=hi or len(r)<=iE: continue
    s=int(r[iSm])
    if s>=thr:
        st=sorted(((int(r[i] or 0),h[6:]) for h,i in sc.items()),reverse=True)[:2]
        print(k, r[iS].strip()[:60].ljust(60), s, st)

import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
seg=0; segs=[collections.Counter() for _ in range(4)]; ss=[0]*4
for k,r in enumerate(rows[2:]):
    if len(r)<=iE: continue
    src=r[iS].strip(); toks=src.split(); op=toks[1] if toks[0].startswith('@') else toks[0]
    if op.startswith("CREDUX") and seg==0: seg=1; print("REDUX at row",k)
    segs[seg][op.split('.')[0]]+=int(r[iE]); ss[seg]+=int(r[iSm])
for s in range(2):
    t=sum(segs[s].values()); print("segment",s,"inst/warp",t/524288,"samples",ss[s])
    print({k:round(v/524288,1) for k,v in segs[s].most_common(14)})

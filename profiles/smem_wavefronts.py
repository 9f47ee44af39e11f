import csv, io, subprocess, sys
rep=sys.argv[1]; steps=float(sys.argv[2])
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur=None;hdr=None;L=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]
    elif len(r)>5 and r[0]=="Line No": hdr={k:i for i,k in enumerate(r)}
    elif hdr and len(r)>10 and r[2]=="-":
        def g(k):
            try: return int(r[hdr[k]])
            except: return 0
        L.append((g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive"), g("L1 Wavefronts Shared Ideal"), cur, r[0], r[1].strip()[:100]))
tot=sum(x[0] for x in L); exc=sum(x[1] for x in L)
print("wavefronts/step %.0f  excessive/step %.0f" % (tot/steps, exc/steps))
for w,e,i,f,ln,src in sorted(L,key=lambda x:-x[0])[:22]:
    print("%7.1f wf/step (excess %6.1f)  %s:%s  %s" % (w/steps, e/steps, f, ln, src))

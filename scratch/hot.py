import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
# may contain several kernels; split on "Kernel Name" rows
i=0; out=[]
while i < len(rows):
    if rows[i] and rows[i][0]=="Kernel Name":
        name=rows[i][1]; hdr=rows[i+1]; j=i+2; body=[]
        while j < len(rows) and not (rows[j] and rows[j][0]=="Kernel Name"): body.append(rows[j]); j+=1
        out.append((name,hdr,body)); i=j
    else: i+=1
name,hdr,body=out[int(sys.argv[2]) if len(sys.argv)>2 else 0]
ie=hdr.index("Instructions Executed"); si=hdr.index("Source"); ss=hdr.index("# Samples")
tot=sum(int(r[ie]) for r in body if len(r)>ie and r[ie].isdigit())
print(name, "total warp-insts", tot, "nlines", len(body))
lo=int(sys.argv[3]) if len(sys.argv)>3 else 0; hi=int(sys.argv[4]) if len(sys.argv)>4 else 10**9
for n,r in enumerate(body):
    if len(r)>ie and r[ie].isdigit() and lo<=n<hi:
        print(f"{n:4d} {int(r[ie]):9d} {100*int(r[ie])/tot:5.2f}% s={r[ss]:>5} {r[si].strip()[:90]}")

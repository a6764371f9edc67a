import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
seq=[(r[ki].split("(")[0], float(r[vi].replace(",",""))/1e3) for r in rows[1:]]
cur=[]; out=[]
for k,v in seq:
    if k=="lg_track_kernel":
        if cur: out.append(cur)
        cur=[]
    cur.append((k,v))
out.append(cur)
sel=[int(x) for x in sys.argv[2].split(",")] if len(sys.argv)>2 else range(len(out)-1)
for i,c in enumerate(out[1:]):
    if i in sel: print(i, " ".join("%s=%.1f"%(k.replace("lg_","").replace("_kernel",""),v) for k,v in c))

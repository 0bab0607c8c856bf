"""Contiguous SASS regions (similar execution count) of one kernel with their share of executed instructions.
usage: python scripts/ncu_regions.py report.ncu-rep [min_share_pct]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu","-i",sys.argv[1],"--page","source","--csv"],capture_output=True,text=True).stdout
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
blk = raw.split('"Kernel Name",')[1]
rows = list(csv.reader(io.StringIO("\n".join(blk.split("\n")[1:]))))
hdr=rows[0]; iS=hdr.index("Source"); iSamp=hdr.index("# Samples"); iInst=hdr.index("Instructions Executed")
data=[(i,r[iS].strip(),int(r[iSamp]),int(r[iInst])) for i,r in enumerate(rows[1:]) if len(r)>iInst and r[iInst].isdigit()]
tot=sum(d[3] for d in data); ts=sum(d[2] for d in data)
reg=[]
for d in data:
    if reg and abs(d[3]-reg[-1][2])<=0.25*max(reg[-1][2],1):
        reg[-1][1]=d[0]; reg[-1][3]+=d[3]; reg[-1][4]+=d[2]; reg[-1][5]+=1
    else:
        reg.append([d[0],d[0],d[3],d[3],d[2],1])
print(f"total {tot:.3e} warp-instructions")
for r in reg:
    if 100*r[3]/tot>=thr:
        print(f"rows {r[0]:4d}-{r[1]:4d} n~{r[2]:.2e} lines {r[5]:3d} inst {100*r[3]/tot:5.1f}% samples {100*r[4]/ts:5.1f}%   first: {data[r[0]][1][:70]}")

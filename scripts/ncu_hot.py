"""Per-SASS hot spots of one kernel from an .ncu-rep (source page): executed-instruction share by opcode
and the top stall-sample addresses.  usage: python scripts/ncu_hot.py report.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = blocks[1 + which]
lines = blk.split("\n")
print("kernel:", lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
iS, iSamp, iInst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_inst = sum(int(r[iInst]) for r in rows[1:] if len(r) > iInst and r[iInst].isdigit())
tot_samp = sum(int(r[iSamp]) for r in rows[1:] if len(r) > iSamp and r[iSamp].isdigit())
by_op = collections.Counter(); samp_op = collections.Counter()
for r in rows[1:]:
    if len(r) <= iInst or not r[iInst].isdigit():
        continue
    toks = r[iS].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    by_op[op] += int(r[iInst]); samp_op[op] += int(r[iSamp])
print(f"total warp-instructions {tot_inst:.3e}, samples {tot_samp}")
for op, n in by_op.most_common(22):
    print(f"  {op:10s} inst {100*n/tot_inst:5.1f}%  samples {100*samp_op[op]/max(tot_samp,1):5.1f}%")

"""Per CUDA source line: share of executed warp instructions and of stall samples, from an .ncu-rep captured with
--import-source on (kernels built with -lineinfo).  usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst = collections.Counter(); samp = collections.Counter(); text = {}
local = collections.Counter()
fname = "?"
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] in ("File Path", "File Name"):
        fname = row[1].split("/")[-1]; continue
    if row[0] in ("Function Name", "Line No") or len(row) < 8:
        continue
    if row[0].isdigit():          # a CUDA source line with its aggregated counters
        key = (fname, int(row[0]))
        text[key] = row[1].strip()
        if row[7].isdigit():
            inst[key] += int(row[7]); samp[key] += int(row[6]) if row[6].isdigit() else 0
        cur = key
    elif row[0] == "" and row[3].strip().startswith(("LDL", "STL")) and row[7].isdigit():
        local[cur] += int(row[7])
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp-instructions {ti:.3e}, samples {ts}, local-memory instructions {sum(local.values()):.3e}")
for key, n in inst.most_common(top):
    print(f"{100*n/ti:5.1f}% inst {100*samp[key]/max(ts,1):5.1f}% samp  LDL/STL {100*local[key]/ti:4.1f}%  {key[0]}:{key[1]:<4d} {text[key][:110]}")
if len(sys.argv) > 3:   # buckets: "name=file:lo-hi,file:lo-hi;name=..."
    print("-- buckets")
    rest = collections.Counter(inst)
    for spec in sys.argv[3].split(";"):
        name, ranges = spec.split("=")
        n = s = l = 0
        for r in ranges.split(","):
            f, lh = r.split(":"); lo, hi = (int(v) for v in lh.split("-"))
            for key in list(inst):
                if key[0] == f and lo <= key[1] <= hi:
                    n += inst[key]; s += samp[key]; l += local[key]; rest.pop(key, None)
        print(f"{100*n/ti:5.1f}% inst {100*s/max(ts,1):5.1f}% samp  LDL/STL {100*l/ti:4.1f}%  {name}")
    print(f"{100*sum(rest.values())/ti:5.1f}% inst  (everything else)")

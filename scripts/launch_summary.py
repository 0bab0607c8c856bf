"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    name = r[iK].split("(")[0]
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>12s} {'avg ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {v:12.3f} {v/cnt[k]:10.4f} {100*v/T:6.2f}%")
print(f"{'TOTAL':70s} {sum(cnt.values()):8d} {T:12.3f}")

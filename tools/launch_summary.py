"""Aggregate an ncu launch list (gpu__time_duration.sum csv) by kernel name.
Usage: python tools/launch_summary.py launches.csv "command line" > profiles/rNN_launches_summary.txt"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)      # -> microseconds
        agg[d["Kernel Name"][:120]][0] += 1
        agg[d["Kernel Name"][:120]][1] += v
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none --csv ;", " ".join(sys.argv[2:]))
print("# cold-cache, serialised launch times: compare SHARES, not absolutes")
print(f"# total device time in list: {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1] / 1e3:10.2f} ms {v[0]:5d} launches {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:8.3f} ms  {k}")

"""Sum an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
acc = collections.OrderedDict()
for r in csv.DictReader(lines):
    if not r["Metric Name"].startswith("gpu__time"):
        continue
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
    k = r["Kernel Name"][:60]
    a = acc.setdefault(k, [0.0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in acc.values())
for k, (us, n) in sorted(acc.items(), key=lambda t: -t[1][0]):
    print(f"{us:10.1f} us {n:5d}x  {100*us/tot:5.1f}%  {k}")
print(f"{tot:10.1f} us total")

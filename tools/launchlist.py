"""Prints an `ncu --metrics gpu__time_duration.sum --csv` launch list with layer names for one trunk forward."""
import csv, sys
names = ["stem", "maxpool"]
for li, nb in zip((1, 2, 3, 4), (3, 4, 6, 3)):
    for b in range(nb):
        names += [f"L{li}b{b}.conv1", f"L{li}b{b}.conv2"] + ([f"L{li}b{b}.down"] if b == 0 else []) + [f"L{li}b{b}.conv3"]
names.append("gap_bn")
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
start = next(i for i, r in enumerate(rows) if "stem" in r["Kernel Name"])
tot = 0.0
for j, row in enumerate(rows[start:start + len(names)]):
    t = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    t = t / 1e3 if u in ("nsecond", "ns") else (t * 1e3 if u in ("msecond", "ms") else t)
    tot += t
    print(f"{names[j]:14s} {row['Kernel Name'][:30]:30s} {t:8.1f} us")
print(f"total {tot:.1f} us over {len(names)} launches")

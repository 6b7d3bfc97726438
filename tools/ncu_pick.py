"""Filters `ncu --page raw --csv` (stdin) down to the metrics matching a regex: prints name, unit, value."""
import csv, re, sys
pat = re.compile(sys.argv[1])
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    names, units = rows[0], rows[1]
    for vals in rows[2:]:
        kn = vals[names.index("Kernel Name")] if "Kernel Name" in names else "?"
        print("# kernel:", kn[:90])
        for n, u, v in zip(names, units, vals):
            if pat.search(n):
                print(f"{n} [{u}] = {v}")

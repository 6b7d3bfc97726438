"""Reads the ncu csv of a `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:dist_gemm`
pass over tools/ncu_retrieval.py and writes profiles/dist_traffic.json: DRAM bytes per dist_gemm_kernel launch (one pass of
config 3), read by bench.py for retrieval.roofline.traffic."""
import csv, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
by_id = {}
for r in csv.DictReader(lines):
    d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u, m = r["Metric Unit"], r["Metric Name"]
    if m.startswith("gpu__time"):
        d["us"] = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
    else:
        d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rows = [d for d in by_id.values() if "dist_gemm" in d["name"]]
per = [d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in rows]
for d, b in zip(rows, per):
    print(f"{d['name'][:40]:40s} {d.get('us', 0):8.1f} us  dram {b / 1e6:8.1f} MB")
if per:
    out = {"dram_bytes_per_pass": sum(per) / len(per), "launches": len(per),
           "algorithmic_bytes_per_pass": (3368 + 15913) * 2048 * 4,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the dist_gemm_kernel launches of "
                     "tools/ncu_retrieval.py (3368 x 15913 x 2048; planes are 2 fp16 planes = the same bytes as fp32 rows)"}
    json.dump(out, open(sys.argv[2], "w"))
    print(out)

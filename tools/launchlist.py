"""Labels an ncu launch list of one trunk forward (csv of `--metrics gpu__time_duration.sum[,dram__bytes_*]`)
with layer names; with DRAM metrics present also prints per-launch traffic and writes the conv total to
profiles/conv_traffic.json (read by bench.py for roofline.traffic)."""
import csv, json, os, sys
FUSED = os.environ.get("CTL_FUSE_SHORTCUT", "1") == "1"  # conv3 + shortcut in one launch (51 launches), else 55
names = ["stem_pack", "stem_pool"]
for li, nb in zip((1, 2, 3, 4), (3, 4, 6, 3)):
    for b in range(nb):
        names += [f"L{li}b{b}.conv1", f"L{li}b{b}.conv2"]
        if b == 0 and not FUSED:
            names += [f"L{li}b{b}.down"]
        names += [f"L{li}b{b}.conv3" + ("+down" if (b == 0 and FUSED) else "")]
names.append("gap_bn")
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
by_id = {}
order = []
for r in rows:
    k = r["ID"]
    if k not in by_id:
        by_id[k] = {"name": r["Kernel Name"]}
        order.append(k)
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("gpu__time"):
        v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
        by_id[k]["us"] = v
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        by_id[k][m] = v * scale
launches = [by_id[k] for k in order]
start = next(i for i, r in enumerate(launches) if "stem_pack" in r["name"])
tot = 0.0
conv_bytes = 0.0
have_dram = False
for j, r in enumerate(launches[start:start + len(names)]):
    tot += r.get("us", 0.0)
    rd, wr = r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum")
    extra = ""
    if rd is not None and wr is not None:
        have_dram = True
        extra = f"  dram rd {rd / 1e6:8.1f} MB  wr {wr / 1e6:8.1f} MB"
        if "conv" in r["name"]:
            conv_bytes += rd + wr
    print(f"{names[j]:14s} {r['name'][:30]:30s} {r.get('us', 0.0):8.1f} us{extra}")
print(f"total {tot:.1f} us over {len(names)} launches")
if have_dram:
    print(f"conv launches: {conv_bytes / 1e9:.3f} GB DRAM traffic per forward")
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump({"dram_bytes_per_step": conv_bytes, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                       f"{sum('conv' in n for n in names)} conv launches of one bs-256 forward (tools/ncu_final.sh)"}, f)

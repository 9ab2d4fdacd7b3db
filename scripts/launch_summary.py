"""Summarise an ncu launch list (--csv, metrics gpu__time_duration.sum + dram bytes): per-kernel launches, total time,
share, DRAM bytes, over the LAST `--steps` captured steps' worth of launches.  usage: python scripts/launch_summary.py file.csv [--tail N]"""
import csv, sys, collections, re
fn = sys.argv[1]
tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 0
rows = []
with open(fn) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
launch = collections.OrderedDict()
for r in rd:
    key = r["ID"]
    d = launch.setdefault(key, {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
L = list(launch.values())
if tail:
    L = L[-tail:]
agg = collections.OrderedDict()
for d in L:
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("ns", 0); a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f us total (serialised, cold-cache)" % (len(L), tot / 1e3))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s x%-4d %9.1f us  %5.1f%%  dram %8.2f MB" % (n, a[0], a[1] / 1e3, 100 * a[1] / tot, a[2] / 1e6))

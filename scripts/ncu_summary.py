"""profiles/ncu_traffic.json from ncu captures of this repo's bench (what bench.py's roofline.traffic reads):
  python scripts/ncu_summary.py <launch list csv of `bench.py --steps K` (metrics gpu__time_duration + dram bytes)> <K> [raw csv of a --set full capture ...]
Per kernel: launches per step, DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), mean duration; "__step__" = the
sum over one step.  The raw csvs (ncu -i x.ncu-rep --page raw --csv) add the tensor-pipe / issue counters of the top kernels."""
import collections, csv, json, re, sys

fn, steps = sys.argv[1], int(sys.argv[2])
with open(fn) as f:
    lines = [l for l in f if l.startswith('"')]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
L = list(launch.values())
# the last `steps` captured-step replays: find the per-step period from the conv kernel count
conv = [i for i, d in enumerate(L) if "conv_pool_tc_kernel" in d["name"]]
per_step_conv = 2
first = conv[-steps * per_step_conv]
# a step starts at the memset/first kernel before its first conv: take launches from the first conv of the window onwards, minus nothing
win = L[first - 8 if first >= 8 else 0:]
agg = collections.OrderedDict()
for d in win:
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("<unnamed>::", "")
    n = re.sub(r"<.*", "", n)
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("ns", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
out = {"source": fn, "steps_in_window": steps, "kernels": {}}
tot_ns = tot_b = 0.0
for n, a in agg.items():
    out["kernels"][n] = {"launches_per_step": a[0] / steps, "dram_bytes_per_launch": a[2] / a[0], "ns_per_launch": a[1] / a[0]}
    tot_ns += a[1]; tot_b += a[2]
out["kernels"]["__step__"] = {"launches_per_step": sum(a[0] for a in agg.values()) / steps, "dram_bytes_per_launch": tot_b / steps, "ns_per_launch": tot_ns / steps}
for raw in sys.argv[3:]:
    with open(raw) as f:
        rows = list(csv.reader(f))
    hdr, vals = rows[0], rows[2:]
    want = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    for v in vals:
        name = v[hdr.index("Kernel Name")]
        k = re.sub(r"<.*|\(.*", "", name.replace("void ", "").replace("<unnamed>::", ""))
        out.setdefault("full_capture", {})[k] = {m: v[hdr.index(m)] for m in want if m in hdr}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
for n, a in sorted(out["kernels"].items(), key=lambda kv: -kv[1]["ns_per_launch"] * kv[1]["launches_per_step"]):
    print("%-40s x%-5.1f %9.1f us  dram %8.2f MB/launch" % (n[:40], a["launches_per_step"], a["ns_per_launch"] / 1e3, a["dram_bytes_per_launch"] / 1e6))

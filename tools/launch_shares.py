"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares_*.txt"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows:
    if r is hdr or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    t = float(r[v].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[u], 1e-6)
    name = r[k].split("(")[0][:78]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"# total {total:.1f} ms over {n} launches (per-launch times under ncu are cold-cache and serialised: compare SHARES)")
print(f"{'kernel':80s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}")
ours = 0.0
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:80s} {c:8d} {t:10.2f} {t / total:7.4f}")
    if "gpfq" in name:
        ours += t
print(f"# libgpfq_b200 kernels: {ours:.1f} ms = {ours / total:.4f} of the step")

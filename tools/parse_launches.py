"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
of bench.py: per-kernel time share and DRAM traffic of ONE train step (between two launches of the step's first kernel).
usage: python tools/parse_launches.py <launches.csv> [first_kernel_substring]"""
import collections
import csv
import re
import sys

fn = sys.argv[1]
first = sys.argv[2] if len(sys.argv) > 2 else "act_embed_kernel"
with open(fn) as f:
    lines = [l for l in f if not l.startswith("==")]
launch = collections.OrderedDict()          # ID -> {name, metric: value}
for r in csv.DictReader(lines):
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    if r["Metric Name"].startswith("gpu__time_duration"):
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)          # -> ns
    elif "bytes" in r["Metric Name"]:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r["Metric Name"]] = v
rows = list(launch.values())
starts = [i for i, r in enumerate(rows) if first in r["name"]]
print("captured", len(rows), "launches; step starts at", starts[:8])
k = min(2, len(starts) - 2)
step = rows[starts[k]:starts[k + 1]]
agg = collections.OrderedDict()
for r in step:
    key = re.sub(r"\(.*", "", r["name"].replace("void ", "")).replace("mfvae::", "")
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += r.get("gpu__time_duration.sum", 0.0)
    a[2] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
tt = sum(a[1] for a in agg.values())
tb = sum(a[2] for a in agg.values())
print("launches in step %d; step total (ncu: serialized, cold cache) %.1f us; DRAM traffic %.1f MB" % (len(step), tt / 1e3, tb / 1e6))
for key, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-52s x%-3d %8.1f us %5.1f%%  dram %8.1f MB  %7.1f GB/s" % (key[:52], c, t / 1e3, 100 * t / tt, b / 1e6, b / max(t, 1e-9)))
gem = [v for k_, v in agg.items() if k_.startswith("gemm_tc_kernel")]
if gem:
    print("gemm_tc_kernel (all instantiations): %d launches, %.1f us (%.1f%% of the step), DRAM traffic %.1f MB per step"
          % (sum(g[0] for g in gem), sum(g[1] for g in gem) / 1e3, 100 * sum(g[1] for g in gem) / tt, sum(g[2] for g in gem) / 1e6))

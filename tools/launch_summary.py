"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: launch_summary.py file.csv [frames]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, per = collections.OrderedDict(), collections.defaultdict(list)
for r in rows[1:]:
    n = re.sub(r"<[^<>]*>$", "", r[ki].split("(")[0].replace("plt::<unnamed>::", "").replace("void ", "").strip())  # template instantiations (k_shade<0>) count as their kernel
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
    per[n].append(v)
build = ("k_world_tris", "k_morton", "k_radix_tree", "k_fit", "k_collapse", "k_sah_costs", "k_ploc_init", "k_ploc_nearest", "k_ploc_merge", "k_ploc_compact")
tot = sum(a[1] for n, a in agg.items() if n.startswith("k_") and n not in build)
if len(sys.argv) <= 2 and "k_raygen" in agg:
    frames = agg["k_raygen"][0]  # one k_raygen launch per frame (untiled frames)
for n, a in agg.items():
    print(f"{n[:60]:60s} launches {a[0]:4d}  ms/frame {a[1] / 1e6 / frames:9.3f}  share {a[1] / tot:.3f}")
for n in ("k_trace", "k_shade", "k_connect", "k_unwind_step"):
    if n in per:
        print(n, "per level (us):", [round(x / 1e3) for x in per[n][:9]])

"""Reads `ncu --page raw --csv` dumps of `--set full` captures and writes profiles/traffic.json: per kernel, the mean over the
captured launches of dram__bytes_read.sum + dram__bytes_write.sum (bench.py's roofline.traffic), with duration, issue-slot
utilisation and warp execution efficiency next to it.  Usage: ncu_traffic.py out.json raw1.csv [raw2.csv ...]"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


acc = collections.defaultdict(lambda: collections.defaultdict(list))
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    rows = [r for r in rows if len(r) > 20]
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}

    def get(r, name, scale=True):
        if name not in col:
            return None
        v = num(r[col[name]])
        if v is None:
            return None
        return v * UNIT.get(units[col[name]], 1.0) if scale else v

    for r in rows[2:]:
        k = re.sub(r"<[^<>]*>$", "", r[col["Kernel Name"]].split("(")[0].split("::")[-1])  # "plt::<unnamed>::k_x<0>(FrameParams)" (ncu may print "unnamed>::k_x") -> "k_x"
        a = acc[k]
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        if rd is not None and wr is not None:
            a["dram"].append(rd + wr)
        for key, name in (("ms", "gpu__time_duration.sum"), ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("regs", "launch__registers_per_thread"),
                          ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                          ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
                          ("local_ld", "smsp__inst_executed_op_local_ld.sum"), ("local_st", "smsp__inst_executed_op_local_st.sum"), ("inst", "smsp__inst_executed.sum")):
            v = get(r, name)
            if v is not None:
                a[key].append(v)

out = {}
for k, a in acc.items():
    n = len(a["ms"]) or 1
    e = {"launches_captured": n}
    if a["dram"]:
        e["dram_bytes_per_launch"] = sum(a["dram"]) / len(a["dram"])
    for key in ("ms", "issue_pct", "threads_per_inst", "regs", "warps_active_pct", "dram_pct", "l1_hit_pct", "l2_hit_pct", "local_ld", "local_st", "inst"):
        if a[key]:
            e[key] = sum(a[key]) / len(a[key])
    out[k] = e
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))

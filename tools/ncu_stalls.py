"""Aggregates the warp-stall samples of an `ncu --page source --csv` dump (SASS view): totals per stall reason, the share of
samples on global/local memory instructions, and the top-N stalled instructions.  Usage: ncu_stalls.py file.csv.gz [kernel-substring] [N]"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    short = b["name"].split("(")[0].split("::")[-1]
    if want not in short or short in seen or len(b["rows"]) < 3:
        continue
    seen.add(short)
    hdr = b["rows"][0]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_")]
    tot = collections.Counter()
    per = []
    inst = 0
    for r in b["rows"][1:]:
        if len(r) < len(hdr):
            continue
        s = int(r[col["# Samples"]] or 0)
        inst += int(r[col["Instructions Executed"]] or 0)
        for n in stall_cols:
            tot[n] += int(r[col[n]] or 0)
        per.append((s, r[col["Source"]].strip(), r[col["Avg. Threads Executed"]], {n: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0)}))
    total = sum(tot.values()) or 1
    print(f"== {short}: {len(per)} SASS instructions, {inst} warp-instructions executed, {total} stall samples")
    print("   " + ", ".join(f"{n[6:]} {v / total:.1%}" for n, v in tot.most_common(8)))
    per.sort(key=lambda x: -x[0])
    for s, src, thr, d in per[:top]:
        why = max(d, key=d.get) if d else "-"
        print(f"   {s / total:6.2%}  thr {thr:>5s}  {why[6:]:14s} {src[:90]}")

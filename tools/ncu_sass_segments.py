"""Lane utilisation along a kernel's SASS from an `ncu --page source --csv` dump: the kernel is cut into blocks of B instructions; per block
the share of executed warp-instructions, the average active threads per instruction and the share of stall samples.
Usage: ncu_sass_segments.py dump.csv.gz [launch index, 1-based] [B]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 40
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows, kern, col, name = [], 0, None, ""
for r in csv.reader(f):
    if r and r[0] == "Kernel Name":
        kern += 1
        if kern == launch:
            name = r[1]
        continue
    if r and r[0] == "Address":
        col = {k: i for i, k in enumerate(r)}
        continue
    if kern != launch or col is None or len(r) < 12:
        continue
    try:
        rows.append((r[1].strip(), int(r[col["Instructions Executed"]]), float(r[col["Avg. Threads Executed"]]), int(r[col["# Samples"]])))
    except ValueError:
        pass
tot, tots = sum(x[1] for x in rows) or 1, sum(x[3] for x in rows) or 1
print(f"{name}: {len(rows)} SASS instructions, {tot} warp-instructions executed, {tots} stall samples, {sum(x[1] * x[2] for x in rows) / tot:.2f} active threads per instruction")
print("first instr | % of executed instructions | active threads | % of stall samples | first instruction of the block | instruction with most samples")
for i in range(0, len(rows), B):
    seg = rows[i:i + B]
    n = sum(x[1] for x in seg)
    if n == 0:
        continue
    print(f"{i:5d} {100 * n / tot:5.1f} % {sum(x[1] * x[2] for x in seg) / n:5.1f} thr {100 * sum(x[3] for x in seg) / tots:5.1f} % | {seg[0][0][:44]:44s} | {max(seg, key=lambda x: x[3])[0][:50]}")

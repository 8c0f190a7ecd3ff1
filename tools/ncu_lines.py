"""Hot source lines of an `ncu --page source --print-source cuda,sass --csv` dump: instructions executed and stall samples per
CUDA source line (summed over the captured launches), top N.  Usage: ncu_lines.py file.csv.gz [N] [kernel-substring]"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
inst, samp, text = collections.Counter(), collections.Counter(), {}
cur_file, cur_fn, col = None, "", None
for r in csv.reader(f):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        col = {n: i for i, n in enumerate(r)}
        continue
    if col is None or want not in cur_fn or not r[0] or not r[0].isdigit() or r[0] == "0":
        continue
    key = (cur_file, int(r[0]))
    try:
        inst[key] += int(r[col["Instructions Executed"]])
        samp[key] += int(r[col["# Samples"]])
    except (ValueError, IndexError):
        continue
    text[key] = r[1].strip()[:110]
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp-instructions {ti}, stall samples {ts}")
byfile_i, byfile_s = collections.Counter(), collections.Counter()
for k, v in inst.items():
    byfile_i[k[0]] += v
for k, v in samp.items():
    byfile_s[k[0]] += v
for fn in sorted(byfile_i, key=byfile_i.get, reverse=True):
    print(f"  {fn:24s} inst {100 * byfile_i[fn] / ti:5.1f}%  samples {100 * byfile_s[fn] / ts:5.1f}%")
print("-- top lines by instructions")
for k, v in inst.most_common(top):
    print(f"{100 * v / ti:5.2f}% inst {100 * samp[k] / ts:5.2f}% smp  {k[0]}:{k[1]:<5d} {text[k]}")
print("-- top lines by stall samples")
for k, v in samp.most_common(top):
    print(f"{100 * v / ts:5.2f}% smp {100 * inst[k] / ti:5.2f}% inst  {k[0]}:{k[1]:<5d} {text[k]}")

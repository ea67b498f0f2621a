"""basic-block view of an `ncu --page source --csv` export: runs of SASS instructions with the same execution count, with their
share of all executed warp instructions, share of stall samples, average active lanes and the most common opcodes.
usage: python tools/sass_blocks.py file.csv.gz [section] [min_share_percent]"""
import csv, gzip, io, sys
from collections import Counter

path = sys.argv[1]; section = int(sys.argv[2]) if len(sys.argv) > 2 else 0; min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
hdr = rows[secs[section] + 1]; data = [r for r in rows[secs[section] + 2:secs[section + 1]] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
tthr = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
print(f"section {section}: {len(data)} SASS instructions, {tot / 1e6:.1f} M warp instructions executed, {tthr / max(tot, 1):.2f} lanes / instruction")
blocks = []; cur = None
for n, r in enumerate(data):
    e = int(r[ix["Instructions Executed"]])
    if cur is None or e != cur["e"]:
        cur = {"e": e, "start": n, "n": 0, "thr": 0, "ops": [], "samples": 0}; blocks.append(cur)
    cur["n"] += 1; cur["thr"] += int(r[ix["Thread Instructions Executed"]]); cur["samples"] += int(r[ix["# Samples"]] or 0)
    src = r[ix["Source"]].split(); cur["ops"].append((src[1] if src and src[0].startswith("@") and len(src) > 1 else src[0] if src else "?").split(".")[0])
alls = sum(b["samples"] for b in blocks)
for b in blocks:
    share = b["e"] * b["n"] / tot
    if 100 * share >= min_share:
        print(f"@{b['start']:5d} n={b['n']:4d} exec={b['e'] / 1e6:8.2f}M share={100 * share:5.1f}% samples={100 * b['samples'] / max(alls, 1):5.1f}% "
              f"lanes={b['thr'] / max(b['e'] * b['n'], 1):5.1f}  {Counter(b['ops']).most_common(8)}")

"""Aggregate an ncu source-page export by source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [kernel-substr] [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
agg = {}
for k, start in enumerate(hdr):
    h = rows[start]
    end = hdr[k + 1] if k + 1 < len(hdr) else len(rows)
    fname = rows[start - 2][1] if start >= 2 else ""
    func = rows[start - 1][1] if start >= 1 else ""
    if want not in func:
        continue
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    for r in rows[start + 1:end]:
        if len(r) > si and r[0].isdigit():
            try:
                s, n = int(r[si]), int(r[ie])
            except ValueError:
                continue
            a = agg.setdefault((fname.split("/")[-1], int(r[0]), r[1][:110]), [0, 0])
            a[0] += s
            a[1] += n
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda t: -t[1][0])[:topn]:
    print(f"{a[0]:6d} {100 * a[0] / tot:5.1f}% inst={a[1]:9d} {k[0]}:{k[1]} {k[2]}")

"""ncu --metrics csv of the RCAN body conv launches -> one line per launch (conv1 = ReLU, conv2 = residual variant)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hd = rows[h]
ki, mi, vi = hd.index("Kernel Name"), hd.index("Metric Name"), hd.index("Metric Value")
d = {}
for r in rows[h + 1:]:
    if len(r) > vi:
        d.setdefault((r[0], "conv2" if "<0, 1" in r[ki] else "conv1"), {})[r[mi].split(".")[0][-28:]] = r[vi]
for k, v in d.items():
    print(k, v)

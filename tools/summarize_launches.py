"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel summary JSON.
usage: python tools/summarize_launches.py launches.csv out.json "what was run" [skip_first_n_launches]"""
import collections, csv, json, re, sys

src, dst, what = sys.argv[1], sys.argv[2], sys.argv[3]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(src, errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n += 1
    if n <= skip:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"\(.*", "", name)
    name = name.replace("mtb::<unnamed>::", "").replace("<unnamed>::", "")[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none --csv", "what": what,
       "total_ms": round(tot, 3), "launches": sum(a[0] for a in agg.values()),
       "kernels": [{"kernel": k, "launches": a[0], "total_ms": round(a[1], 3), "avg_ms": round(a[1] / a[0], 4),
                    "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]]}
json.dump(out, open(dst, "w"), indent=1)
for k in out["kernels"][:14]:
    print(f'{k["total_ms"]:9.3f} ms {k["launches"]:5d} x {k["avg_ms"]:8.4f}  {k["share"]*100:5.1f}%  {k["kernel"]}')
print("total", out["total_ms"], "ms in", out["launches"], "launches")

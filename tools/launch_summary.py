"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time share per kernel name."""
import csv, sys, collections
path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1.0}.get(r[iu], 1.0)
    t = tot.setdefault(r[ik], [0.0, 0])
    t[0] += v; t[1] += 1
total = sum(t[0] for t in tot.values())
print("# %s" % title)
print("# gpu__time_duration.sum per kernel over the whole process (cold-cache, serialised: compare SHARES); total=%.3f ms" % (total / 1e6))
for k, (v, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
    print("%8.4f%%  %14.1f ns  x%-4d %s" % (100 * v / total, v, c, k[:130]))

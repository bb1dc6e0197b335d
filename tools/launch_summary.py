"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0]
    tot[name] += v; cnt[name] += 1
all_ms = sum(tot.values())
print("# launch list summary: %s" % (sys.argv[2] if len(sys.argv) > 2 else path))
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
print("# %-78s %6s %10s %6s" % ("kernel", "count", "total_ms", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-80s %6d %10.3f %5.1f%%" % (k[:80], cnt[k], v, 100 * v / all_ms))

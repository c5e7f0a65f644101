"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
if "--scan" in sys.argv:   # keep only the n-th scan (1-based); a scan starts at live_scan_kernel
    want = int(sys.argv[sys.argv.index("--scan") + 1])
    starts = [i for i, r in enumerate(body) if "live_scan_kernel" in r[ki]]
    lo = starts[want - 1]
    hi = starts[want] if want < len(starts) else len(body)
    body = body[lo:hi]
    print("scan %d of %d: launches %d..%d" % (want, len(starts), lo, hi))
agg = collections.OrderedDict()
for r in body:
    name = r[ki].split("(")[0].replace("void ", "").replace("mht::", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-34s %6s %12s %12s %7s" % ("kernel", "n", "total ms", "avg us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %6d %12.3f %12.1f %6.1f%%" % (k, a[0], a[1] / 1e3, a[1] / a[0], 100 * a[1] / tot))
print("%-34s %6d %12.3f" % ("TOTAL", sum(a[0] for a in agg.values()), tot / 1e3))

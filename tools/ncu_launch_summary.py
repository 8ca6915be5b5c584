"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv ...`):
launches, total time and share of the kernel time per kernel name.
usage: python tools/ncu_launch_summary.py launches.csv [title line ...]"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
tot = collections.Counter()
cnt = collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
for r in rd:
    if len(r) <= max(ik, iv, iu) or r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    tot[name] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    cnt[name] += 1
all_ms = sum(tot.values())
for t in sys.argv[2:]:
    print(t)
print("kernel, launches, total ms, share of kernel time")
for name, ms in tot.most_common():
    print("%s, %d, %.2f, %.2f %%" % (name, cnt[name], ms, 100.0 * ms / all_ms))
print("all kernels, %d, %.2f, 100 %%" % (sum(cnt.values()), all_ms))

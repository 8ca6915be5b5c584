"""Summarise an ncu report's source page: stall samples per source line (needs -lineinfo) and per opcode.
usage: python tools/ncu_src_summary.py report.ncu-rep [top]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] if False else
                     ["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) > 10 and r[0].startswith("0x")]
s = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(float(r[s] or 0) for r in data)
ops = collections.Counter()
for r in data:
    toks = r[1].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    ops[op] += float(r[s] or 0)
print("total samples", tot)
for k, v in ops.most_common(14):
    print("  %-22s %5.1f%%" % (k, 100 * v / tot))
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
agg = collections.Counter()
for r in data:
    for n, i in idx.items():
        agg[n] += float(r[i] or 0)
print("stall reasons:", ", ".join("%s=%.1f%%" % (k, 100 * v / tot) for k, v in agg.most_common(8)))
# cumulative by address region (every 5% of instructions)
n = len(data); chunk = max(1, n // 20); 
print("samples by instruction range (of %d instrs):" % n)
for a in range(0, n, chunk):
    seg = data[a:a + chunk]; v = sum(float(r[s] or 0) for r in seg)
    if v / tot > 0.02:
        print("  [%5d..%5d) %5.1f%%  first: %s" % (a, a + len(seg), 100 * v / tot, seg[0][1].strip()[:60]))

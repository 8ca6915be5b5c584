"""Stall samples per CUDA source line from an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_line_summary.py report.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
lines = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    s = hdr.index("Warp Stall Sampling (All Samples)")
    if r[2] not in ("", "-"):  # a SASS row (has an address): skip, the line row carries the aggregate
        continue
    try:
        v = float(r[s] or 0)
    except ValueError:
        continue
    if v <= 0:
        continue
    reasons = {h: float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    key = (cur_file, int(r[0]))
    ent = lines.setdefault(key, {"src": r[1].strip(), "v": 0.0, "reasons": collections.Counter(), "inst": 0.0})
    ent["v"] += v
    ent["reasons"].update(reasons)
    try:
        ent["inst"] += float(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        pass
tot = sum(e["v"] for e in lines.values())
print("total stall samples: %.0f over %d source lines" % (tot, len(lines)))
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["v"])[:top]:
    rs = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(1.0, sum(e["reasons"].values())))
                   for k, v in e["reasons"].most_common(3))
    print("%5.1f%%  %-16s:%-4d inst=%-9.0f %-62s | %s" % (100 * e["v"] / tot, f, ln, e["inst"], e["src"][:62], rs))

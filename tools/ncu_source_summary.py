"""Summarise an `ncu --page source --csv` export: stall samples per opcode class and per code region.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.defaultdict(lambda: collections.Counter())
total = collections.Counter()
region = collections.defaultdict(lambda: collections.Counter())
cur = "prologue"
n_bar = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    m = re.match(r"\s*(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", src)
    op = m.group(1) if m else "?"
    samples = int(r[col["# Samples"]] or 0)
    execd = int(r[col["Instructions Executed"]] or 0)
    by_op[op]["samples"] += samples
    by_op[op]["exec"] += execd
    by_op[op]["n"] += 1
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        by_op[op][s] += v
        total[s] += v
        region[cur][s] += v
    region[cur]["samples"] += samples
    region[cur]["exec"] += execd
    if op == "BAR":
        n_bar += 1
        cur = "after_bar%d" % n_bar
tot_samples = sum(v["samples"] for v in by_op.values())
print("total samples", tot_samples)
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(1, sum(total.values()))) for k, v in total.most_common(8)))
print("\nper opcode (top by samples):")
for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["samples"])[:22]:
    top = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:3]
    print("  %-10s static=%4d exec=%10d samples=%7d (%.1f%%)  %s" % (
        op, c["n"], c["exec"], c["samples"], 100.0 * c["samples"] / tot_samples,
        " ".join("%s:%d" % (n, v) for v, n in top)))
print("\nper region:")
for k, c in region.items():
    top = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:4]
    print("  %-12s exec=%10d samples=%7d (%.1f%%) %s" % (k, c["exec"], c["samples"], 100.0 * c["samples"] / tot_samples,
                                                     " ".join("%s:%d" % (n, v) for v, n in top)))

"""Summarise an `ncu --page source --csv` export by barrier-delimited segment (development aid)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
segs = []; cur = collections.Counter(); cur_ops = collections.Counter(); start = None
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def flush(end):
    global cur, cur_ops, start
    if cur["n"]:
        segs.append((start, end, cur, cur_ops))
    cur = collections.Counter(); cur_ops = collections.Counter(); start = None
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    if start is None: start = r[ix["Address"]]
    cur["n"] += 1
    cur["samples"] += int(r[ix["# Samples"]] or 0)
    cur["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for s in stall_cols: cur[s] += int(r[ix[s]] or 0)
    cur_ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    if op.startswith("BAR") or op.startswith("SYNCS"): flush(r[ix["Address"]])
flush("end")
tot = sum(s[2]["samples"] for s in segs)
for a, b, c, ops in segs:
    if c["samples"] < tot * 0.005: continue
    top = ", ".join("%s=%d" % (k, v) for k, v in ops.most_common(8))
    st = ", ".join("%s=%.0f%%" % (k[6:], 100 * c[k] / max(c["samples"], 1)) for k in sorted(stall_cols, key=lambda k: -c[k])[:5])
    print("%s..%s n=%d samples=%.1f%% inst=%.2e | %s | %s" % (a[-5:], b[-5:], c["n"], 100 * c["samples"] / tot, c["inst"], top, st))

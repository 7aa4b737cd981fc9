"""Windowed stall summary of an `ncu --page source --csv` export (development aid)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 120
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
for i in range(0, len(body), W):
    seg = body[i:i + W]
    smp = sum(int(r[ix["# Samples"]] or 0) for r in seg)
    if smp < tot * 0.01:
        continue
    st = collections.Counter(); ops = collections.Counter()
    for r in seg:
        for s in stall_cols:
            st[s] += int(r[ix[s]] or 0)
        t = r[ix["Source"]].split(); op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    print(i, seg[0][ix["Address"]][-5:], "samples %.1f%%" % (100 * smp / tot), "inst %.1e" % inst,
          ", ".join("%s=%.0f%%" % (k[6:], 100 * v / max(smp, 1)) for k, v in st.most_common(4)), "|",
          ", ".join("%s%d" % (k, v) for k, v in ops.most_common(4)))

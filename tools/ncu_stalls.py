"""Warp-stall summary of one kernel from an .ncu-rep (`--set full --import-source on`): totals per stall reason, the
producer main loop of dla_loglik_i8p_kernel (located by its SASS markers: from the first MUFU.RCP64H after the last
LDTM to the mbarrier arrive that follows the last STS.U8), and the most-sampled instructions.
usage: python tools/ncu_stalls.py report.ncu-rep out.txt"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
src = [r[ix["Source"]].strip() for r in data]
n = [int(r[ix["# Samples"]]) for r in data]
def summary(lo, hi_, name, f):
    tot = sum(n[lo:hi_])
    s = sorted(((sum(int(r[ix[k]]) for r in data[lo:hi_]), k[6:]) for k in stalls), reverse=True)
    f.write("%s: SASS rows %d..%d, %d samples (%.1f %% of the kernel)\n" % (name, lo, hi_, tot, 100.0 * tot / max(sum(n), 1)))
    for v, k in s[:10]:
        if v: f.write("    %-20s %8d  %5.1f %%\n" % (k, v, 100.0 * v / max(tot, 1)))
ldtm = [i for i, x in enumerate(src) if "LDTM" in x]
sts8 = [i for i, x in enumerate(src) if x.startswith("STS.U8")]
with open(out, "w") as f:
    f.write("# %s\n# kernel: %s\n" % (" ".join(sys.argv), rows[0][1] if len(rows[0]) > 1 else "?"))
    summary(0, len(data), "whole kernel", f)
    if ldtm and sts8:
        lo = next(i for i in range(ldtm[-1], len(src)) if "MUFU.RCP64H" in src[i])
        hi2 = next(i for i in range(sts8[-1], len(src)) if "SYNCS.ARRIVE" in src[i]) + 1
        summary(lo, hi2, "producer main loop (approx.)", f)
    f.write("most-sampled instructions:\n")
    for i in sorted(range(len(data)), key=lambda i: -n[i])[:25]:
        top = sorted(((int(data[i][ix[k]]), k[6:]) for k in stalls), reverse=True)[0]
        f.write("    row %5d  %-60s %7d  (%s %d)\n" % (i, src[i][:60], n[i], top[1], top[0]))
print(open(out).read())

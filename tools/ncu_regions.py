"""Where the warp-stall samples of the persistent INT8 kernel's producer loop go, by code region (development aid).
Reads an .ncu-rep captured with `--set full --import-source on`; regions are found from SASS markers.
usage: python tools/ncu_regions.py report.ncu-rep [rows_per_bucket]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
n = [int(r[ix["# Samples"]]) for r in data]
ex = [int(r[ix["Instructions Executed"]]) for r in data]
src = [r[ix["Source"]].strip() for r in data]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(n)
loop_ex = max(ex[i] for i in range(len(ex)) if src[i].startswith("STS.U8")) if any(s.startswith("STS.U8") for s in src) else max(ex)
print("kernel samples %d; producer main loop executes %d warp-iterations" % (tot, loop_ex))
KEYS = ("MUFU.RCP64H", "LDG.E.64.CONSTANT", "LDG.E.64", "LDG.E.128", "STS.U8", "STS", "LDS.64", "STS.64", "SHFL.IDX", "CALL.REL.NOINC", "DFMA", "DMUL", "DADD",
        "SYNCS.PHASECHK.TRANS64.TRYWAIT", "SYNCS.ARRIVE.TRANS64", "FENCE.VIEW.ASYNC.S", "VOTE.ANY", "UTCIMMA", "LDTM")
for lo in range(0, len(data), step):
    seg = range(lo, min(lo + step, len(data)))
    sm = sum(n[i] for i in seg)
    if sm < tot * 0.002:
        continue
    ops = collections.Counter((s.split()[1] if s.startswith("@") else s.split()[0]) if s else "" for s in (src[i] for i in seg))
    top = sorted(((sum(int(data[i][ix[k]]) for i in seg), k[6:]) for k in stalls), reverse=True)[:3]
    print("%5d %8d %5.1f%%  ex<=%-11d %-50s %s" % (lo, sm, 100.0 * sm / tot, max(ex[i] for i in seg),
                                                   " ".join("%s:%d" % (k, v) for v, k in top if v), {k: ops[k] for k in KEYS if ops.get(k)}))

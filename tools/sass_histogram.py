"""Opcode histogram of the shipped kernels of libgpdla.so (cuobjdump -sass) -- evidence that the hot path is sm_100a code:
tcgen05 MMAs (UTCIMMA), TMEM loads (LDTM), TMA / DSMEM bulk copies (UBLKCP), mbarriers (SYNCS), FP64 tensor MMAs (DMMA).
usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gp_dla_detection_b200", "libgpdla.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "PRMT", "RED", "ATOM",
       "UCGABAR", "USETMAXREG", "FENCE", "CALL", "BAR")
print("# cuobjdump -sass gp_dla_detection_b200/libgpdla.so : instructions per kernel, selected opcode families (prefix match)")
print("# arch: %s" % ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
for k, h in hist.items():
    tot = sum(h.values())
    fam = collections.OrderedDict((f, sum(v for o, v in h.items() if o.startswith(f))) for f in KEY)
    print("%-110s %7d instr  %s" % (k[:110], tot, " ".join("%s=%d" % (f, n) for f, n in fam.items() if n)))

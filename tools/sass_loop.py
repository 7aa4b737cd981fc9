"""Instruction mix of the producer main loop of dla_loglik_i8p_kernel from `cuobjdump -sass` (development aid): the loop
is located by its fence.proxy.async (FENCE.VIEW.ASYNC) and the backward branch that follows it.
usage: python tools/sass_loop.py [libgpdla.so] [mangled kernel name]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gp_dla_detection_b200/libgpdla.so"
fun = sys.argv[2] if len(sys.argv) > 2 else "_ZN5gpdla2i821dla_loglik_i8p_kernelILi20ELi6ELi3ELi0EEEvNS_10LoglikArgsENS0_6I8ArgsEii"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = []
for l in txt.splitlines():
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions")
for i in [i for i, (a, t) in enumerate(ins) if 'FENCE.VIEW.ASYNC' in t]:
    for j in range(i, min(i + 300, len(ins))):
        m = re.search(r'BRA.*0x([0-9a-f]+)', ins[j][1])
        if m and int(m.group(1), 16) < ins[j][0] - 2000:
            k = [x for x, (a, _) in enumerate(ins) if a == int(m.group(1), 16)][0]
            c = collections.Counter()
            for a, t in ins[k:j + 1]:
                op = t.split()[1] if t.startswith('@') else t.split()[0]
                c[op.split('.')[0]] += 1
            print("loop: %d instructions (static, both table and direct paths)" % (j - k + 1))
            print(c.most_common())
            open('/tmp/loop.txt', 'w').write("\n".join("%d %s" % (n, t) for n, (a, t) in enumerate(ins[k:j + 1])))
            break

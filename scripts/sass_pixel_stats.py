"""Per-pixel SASS statistics of a head_kernel instantiation: instructions between consecutive MUFU.RCP (one per pixel).
   cuobjdump -sass -fun <mangled> build/obj/head.o | python scripts/sass_pixel_stats.py [dump_index]"""
import re, sys, collections
lines = [l for l in sys.stdin if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l)]
ops = []
for l in lines:
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_\.]+)(.*?);', l)
    ops.append((int(m.group(1), 16), (m.group(2) or '').strip(), m.group(3), m.group(4)))
print("total SASS instructions:", len(ops))
c = collections.Counter(o[2].split('.')[0] for o in ops)
print("local memory:", {k: v for k, v in c.items() if k in ("STL", "LDL")})
rcps = [i for i, o in enumerate(ops) if o[2].startswith('MUFU.RCP')]
print("gaps between MUFU.RCP:", [rcps[i + 1] - rcps[i] for i in range(len(rcps) - 1)])
if len(sys.argv) > 1:
    k = int(sys.argv[1])
    a, b = rcps[k], rcps[k + 1]
    cc = collections.Counter(o[2].split('.')[0] for o in ops[a:b])
    print(sorted(cc.items(), key=lambda x: -x[1]))
    for o in ops[a:b]:
        print(hex(o[0]), o[1], o[2], o[3])

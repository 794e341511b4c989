"""Development: static view of a kernel's SASS — backward branches (loops) with their length and opcode mix.
    cuobjdump -sass -fun <mangled> obj.o > k.sass ; python tools/sass_loops.py k.sass [lo hi]"""
import collections, re, sys
ins = []
for line in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
def op(t):
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', t)
    return m.group(2).split('.')[0] if m else '?'
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    c = collections.Counter(op(t) for a, t in ins if lo <= a <= hi)
    n = sum(c.values())
    print(f"[{lo:#x}, {hi:#x}] {n} instructions")
    for k, v in c.most_common(40):
        print(f"  {k:10s} {v:5d} {100*v/n:5.1f}%")
else:
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(?:UR\d+,\s*)?(0x[0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            print(f"loop {tgt:#x}..{a:#x}  {(a - tgt)//16 + 1:5d} instr   {t}")

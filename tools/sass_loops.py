"""List the loops of one kernel in a cuobjdump -sass dump with their instruction mix.
Usage: cuobjdump -sass lib.so > all.sass; python tools/sass_loops.py all.sass <mangled-name-substring>"""
import re
import sys

txt = open(sys.argv[1]).read()
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [f for f in funcs if sys.argv[2] in f.split("\n")[0]][0]
ins = []
for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\d\s+)?)([A-Z0-9_.]+)([^;]*);", f):
    ins.append((int(m.group(1), 16), m.group(2).strip(), m.group(3), m.group(4).strip()))
addr = {a: i for i, (a, _, _, _) in enumerate(ins)}
loops = []
for i, (a, p, op, args) in enumerate(ins):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", args)
        if m:
            t = int(m.group(1), 16)
            if t < a and t in addr:
                loops.append((addr[t], i))
print(f.split("\n")[0], len(ins), "instructions")
for s, e in loops:
    ops = {}
    for _, _, op, _ in ins[s : e + 1]:
        o = op.split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    print(hex(ins[s][0]), hex(ins[e][0]), e - s + 1, sorted(ops.items(), key=lambda x: -x[1])[:16])

"""Opcode histogram of a kernel's SASS (whole kernel and its largest backward-branch loop) - runs without a GPU.
usage: sass_mix.py LIB.so KERNEL_NAME_FRAGMENT"""
import collections, re, subprocess, sys
lib, frag = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if frag in f.split("\n")[0])
ops = []
for l in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops.append((int(m.group(1), 16), m.group(3), l))
loops = []
for addr, op, l in ops:
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
print(body.split("\n")[0], "instructions", len(ops), "loops", [(hex(a), hex(b)) for a, b in loops])
if loops:
    lo, hi = max(loops, key=lambda x: x[1] - x[0])
    c = collections.Counter(op.split(".")[0] if len(sys.argv) < 4 else op for a, op, l in ops if lo <= a <= hi)
    n = sum(c.values())
    print("largest loop", hex(lo), hex(hi), "instructions", n)
    print(", ".join(f"{k} {v}" for k, v in c.most_common(45)))

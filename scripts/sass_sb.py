"""Decode SASS control words (stall / yield / write-barrier / read-barrier / wait-mask) of one kernel.
Usage: python scripts/sass_sb.py <lib.so> <mangled-name-substring> [--loads]
Layout (Volta..Blackwell 128-bit encodings): bits 105-108 stall, 109 yield, 110-112 write barrier,
113-115 read barrier, 116-121 wait mask, 122-125 reuse."""
import re, subprocess, sys
lib, name = sys.argv[1], sys.argv[2]
only = "--loads" in sys.argv
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if "Function :" in l and name in l)
ins = []
i = start + 1
while i < len(txt) and "Function :" not in txt[i]:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", txt[i])
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            c = hi >> 41
            stall, yld, wb, rb, wait = c & 15, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 63
            ins.append((int(m.group(1), 16), m.group(2).strip(), stall, yld, wb, rb, wait))
            i += 2
            continue
    i += 1
for n, (addr, s, stall, yld, wb, rb, wait) in enumerate(ins):
    w = ",".join(str(b) for b in range(6) if wait >> b & 1)
    tag = f"{n:4d} {addr:05x} st{stall:2d} {'Y' if yld else ' '} W{wb if wb != 7 else '-'} R{rb if rb != 7 else '-'} wait[{w:11s}] {s}"
    if not only or wb != 7 or wait or "BRA" in s or "STG" in s:
        print(tag)

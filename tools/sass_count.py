#!/usr/bin/env python
"""Static opcode count of a kernel of libcracks_b200.so: whole kernel and the largest backward-branch loop body
(the Gauss-plane loop of the tiled apply kernels, executed 3 times per cell).
usage: python tools/sass_count.py <substring of the mangled kernel name> [library]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "cracks_b200/libcracks_b200.so"
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", names)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = []
    for l in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)([^;]*);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(3).split(".")[0], m.group(4)))
    # largest backward branch
    best = (0, 0)
    for a, op, rest in ins:
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < a and a - int(t.group(1), 16) > best[1] - best[0]:
                best = (int(t.group(1), 16), a)
    loop = collections.Counter(op for a, op, _ in ins if best[0] <= a <= best[1])
    allc = collections.Counter(op for a, op, _ in ins)
    fp64 = lambda c: sum(c[o] for o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    fp32 = lambda c: sum(c[o] for o in ("FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2"))
    print(name[:110])
    print("  total %d instr, loop [%x,%x] %d instr: FP64 %d (DFMA %d DADD %d DMUL %d) FP32 %d  LDS %d STS %d LDG %d | outside loop FP64 %d FP32 %d"
          % (len(ins), best[0], best[1], sum(loop.values()), fp64(loop), loop["DFMA"], loop["DADD"], loop["DMUL"], fp32(loop),
             loop["LDS"], loop["STS"], loop["LDG"], fp64(allc) - fp64(loop), fp32(allc) - fp32(loop)))

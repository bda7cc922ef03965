#!/usr/bin/env python3
"""Groups tools/ncu_lines.py output by the enclosing function of the elector_b200/csrc source file of each line.
  python tools/ncu_lines.py sass.csv dis.txt KERNEL 1000 | python tools/ncu_funcs.py"""
import bisect
import os
import re
import sys
from collections import defaultdict

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "elector_b200", "csrc")
_marks = {}


def marks_of(fname):
    if fname not in _marks:
        path = os.path.join(CSRC, fname)
        if not os.path.exists(path):
            _marks[fname] = None
        else:
            src = open(path).read().split("\n")
            m = [(i + 1, l.strip()[:70]) for i, l in enumerate(src)
                 if re.match(r"\s*(template|EL_HDN|EL_HD|SP_HD|static EL_HD|__global__|__device__|__host__ __device__)", l) and "(" in l
                 and not l.strip().startswith("template <")]
            _marks[fname] = ([x[0] for x in m], m)
    return _marks[fname]


I, S = defaultdict(float), defaultdict(float)
for l in sys.stdin:
    m = re.match(r"\s*([0-9.]+)% inst\s+([0-9.]+)% stall\s+(\S+):(\d+)", l)
    if not m:
        if l.startswith("opcode") or l.startswith("total"):
            print(l.strip())
        continue
    p, s, f, n = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    mk = marks_of(f)
    if mk is None:
        g = f
    else:
        k = bisect.bisect_right(mk[0], n) - 1
        g = f + ": " + (mk[1][k][1] if k >= 0 else "top")
    I[g] += p
    S[g] += s
for g in sorted(I, key=lambda g: -S[g]):
    print("%6.2f%% inst %6.2f%% stall  %s" % (I[g], S[g], g))

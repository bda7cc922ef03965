#!/usr/bin/env python3
"""Attribute an ncu capture to source lines without a GPU.
  ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > sass.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c capi.sm_100a.cubin > dis.txt
  python tools/ncu_lines.py sass.csv dis.txt KERNEL_SUBSTRING [top]
Prints, per source line of the kernel: share of executed warp instructions, share of stall samples."""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# address -> (file, line) from nvdisasm
addr2line, cur, inside = {}, None, False
for ln in open(dis, errors="replace"):
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ia, ie, ismp, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
base = None
per, ops = defaultdict(lambda: [0, 0]), defaultdict(int)
tot = tots = 0
for r in rows[hi + 1:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    n, s = int(r[ie] or 0), int(r[ismp] or 0)
    key = addr2line.get(a - base, (("?", 0), ""))[0]
    per[key][0] += n
    per[key][1] += s
    ops[r[isrc].split()[0].split(".")[0] if not r[isrc].startswith("@") else r[isrc].split()[1].split(".")[0]] += n
    tot += n
    tots += s
print("total warp instructions %d, samples %d" % (tot, tots))
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% inst %6.2f%% stall  %s:%d" % (100.0 * v[0] / tot, 100.0 * v[1] / max(1, tots), k[0], k[1]))
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]))

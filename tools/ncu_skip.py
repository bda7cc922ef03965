#!/usr/bin/env python3
"""From the ncu launch list of `pipe_driver PREFIX 2` (two calls; the first one grows the scratch pools and runs some segments
twice): how many launches matching REGEX precede the LAST call and how many belong to it -> "SKIP COUNT" for
`ncu -k regex:REGEX --launch-skip SKIP -c COUNT` of the same command.
  python tools/ncu_skip.py LAUNCHES.csv REGEX"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
pat = re.compile(sys.argv[2])
hdr, names = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        names.append(dict(zip(hdr, r))["Kernel Name"])
last = max(i for i, n in enumerate(names) if n.startswith("init_call_kernel"))
skip = sum(1 for n in names[:last] if pat.search(n))
count = sum(1 for n in names[last:] if pat.search(n))
print(skip, count)

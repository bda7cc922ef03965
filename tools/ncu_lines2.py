#!/usr/bin/env python3
"""Per source line of one kernel of an ncu source page: share of warp instructions, active threads per instruction, share of
stall samples and their split into barrier / long scoreboard / short scoreboard / lg.
  ncu -i X.ncu-rep --page source --csv > src.csv; nvdisasm -g -c capi.sm_100a.cubin > dis.txt
  python tools/ncu_lines2.py src.csv dis.txt KERNEL_SUBSTRING"""
import csv, re, sys
from collections import defaultdict
src, dis, kern = sys.argv[1:4]
addr2line = {}; cur = None; inside = False
for ln in open(dis, errors='replace'):
    if ln.startswith('//---') and '.text.' in ln: inside = kern in ln
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
h = rows[hi]
ia, ie, it, ismp = h.index('Address'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
cols = [h.index(c) for c in ('stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_lg', 'stall_wait', 'stall_branch_resolving')]
per = defaultdict(lambda: [0] * 9); base = None; T = S = 0
for r in rows[hi + 1:]:
    try: a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    except ValueError: continue
    if base is None: base = a
    v = per[addr2line.get(a - base)]
    v[0] += int(r[ie] or 0); v[1] += int(r[it] or 0); v[2] += int(r[ismp] or 0)
    for j, c in enumerate(cols): v[3 + j] += int(r[c] or 0)
    T += int(r[ie] or 0); S += int(r[ismp] or 0)
print('total warp instructions %d, samples %d' % (T, S))
print('line                          inst%  thr/inst  samples% | barrier long_sb short_sb lg wait branch')
for k, v in sorted(per.items(), key=lambda kv: (kv[0] or ('', 0))):
    if v[0] * 100 / T > 0.3 or v[2] * 100 / S > 0.5:
        print('%-28s %6.2f %6.1f %8.2f | ' % ('%s:%d' % k if k else '?', v[0] * 100 / T, v[1] / max(1, v[0]), v[2] * 100 / S) + ' '.join('%5.2f' % (x * 100 / S) for x in v[3:]))

#!/bin/bash
# SASS evidence of the hot loops (no GPU needed): for the bulk DP kernels of libelector_poa.so the packed halfword mnemonics,
# the instruction mix of the whole kernel, and the innermost band loop of the dual kernel (the backward branch with the most
# VIADDMNMX.S16x2 inside) with its instructions per iteration and per pair of cells.
#   bash tools/sass_excerpt.sh > profiles/rNN_sass_hot_loops.txt
LIB=${1:-elector_b200/libelector_poa.so}
cuobjdump -sass $LIB > /tmp/all.sass
echo "# $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
for k in 7Phase2DELi 7Phase2LELi 7Phase1PELi; do
  awk -v k="$k" '/Function : /{f = index($0, k) > 0 && index($0, "poa_dp") > 0} f' /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4,5})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/k.sass
  echo; echo "== kernel *${k}*: $(wc -l < /tmp/k.sass) instructions"
  echo "packed halfword / byte-permute instructions:"; grep -oE "VIMNMX\.[SU]16x2|VIADDMNMX\.S16x2|PRMT|UTMALDG|UBLKCP|LDGSTS" /tmp/k.sass | sort | uniq -c | sort -rn | sed 's/^/   /'
  echo "opcode mix (top 12):"; awk '{print $2}' /tmp/k.sass | sed 's/^@!*U*P[0-9T]*$//' | awk 'NF' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -12 | sed 's/^/   /'
done
# the band loops: per kernel the innermost loop (backward branch) that holds at least 6 VIADDMNMX.S16x2 (one per register = per
# pair of cells of an update; R = 6, 7 or 8 registers, some loops are unrolled twice by the compiler)
for k in 7Phase2DELi 7Phase2LELi 7Phase1PELi; do
  awk -v k="$k" '/Function : /{f = index($0, k) > 0 && index($0, "poa_dp") > 0} f' /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4,5})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/d.sass
  KNAME=$k python3 - <<'PY'
import os
import re
L = [l.rstrip("\n") for l in open("/tmp/d.sass")]
addr = {int(l.split()[0], 16): i for i, l in enumerate(L)}
best = None
for i, l in enumerate(L):
    m = re.search(r"BRA\s+(0x[0-9a-f]+)", l)
    if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
        j = addr[int(m.group(1), 16)]
        n = sum("VIADDMNMX" in x for x in L[j:i + 1])
        if n >= 6 and (best is None or i - j < best[2] - best[1]):
            best = (n, j, i)
n, j, i = best
body = L[j:i + 1]
print("\n== innermost band loop of the kernel *%s*: SASS %s .. %s" % (os.environ["KNAME"], L[j].split()[0], L[i].split()[0]))
print("instructions in the loop body: %d; VIADDMNMX.S16x2: %d = pairs of cells per trip" % (len(body), n))
skip = None   # a block skipped by a forward branch in the common trip (the dual kernel: the end of a bubble)
for k, l in enumerate(body):
    m = re.search(r"@!?P\d+\s+BRA\s+(0x[0-9a-f]+)", l)
    if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] - j > k + 60:
        skip = (k, addr[int(m.group(1), 16)] - j)
        break
common = len(body) - (skip[1] - skip[0] - 1) if skip else len(body)
print("common trip%s: %d instructions = %.1f per pair of cells, %.1f per cell" % (" (the block of %d instructions that closes a bubble skipped)" % (skip[1] - skip[0] - 1) if skip else "", common, common / n, common / (2 * n)))
print("--- loop body ---")
for l in body:
    print("  " + l)
PY
done

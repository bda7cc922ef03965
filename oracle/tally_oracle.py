"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the merge + tally half of the hot path.

  merge_windows()   src/split/Donatello.cpp:13-31 (clean_msa) and :50-84 (per-read concat)
  tally_read()      elector/computeStats.py, integer part of the per-read tally:
      left/right gaps      nbLeftGaps :61-77, nbRightGaps :82-98
      gaps + extension     gapsAndExtensions :472-498
      gap stretches        findGapStretches :104-189
      existing positions   getCorrectedPositions :712-752 (no clipping: -simulator real only)
      per-column counters  getTPFNFP :399-440, indels :291-328, getCorrectionAtEachPosition :371-393
Parity status: PINNED -- tests/test_tally_oracle.py compares every function here with
outputs of the reference's own Python functions (tests/golden/tally_*.json.gz, generated
by oracle/make_golden_tally.py importing /root/reference/elector/computeStats.py) and of
the compiled reference Donatello.
"""
THRESH = 5
THRESH2 = 20

FIELDS = ["TP", "FP", "FN", "cor", "uncor", "uncorCor", "uncorUncor", "insC", "delC", "subsC",
          "insU", "delU", "subsU", "GCref", "GCcor", "lenRef", "lenCor", "lenUnc", "gapsLeft",
          "gapsRight", "missing", "extended", "ncols", "assessed"]


def nb_left_gaps(s):
    gaps = nt = total = 0
    i = 0
    while i < len(s) and nt <= THRESH:
        if s[i] == ".":
            gaps += 1
            nt = 0
        else:
            if gaps >= THRESH:
                total = i
            gaps = 0
            nt += 1
        i += 1
    return total


def nb_right_gaps(s):
    gaps = nt = total = 0
    i = len(s) - 1
    while i >= 0 and nt <= THRESH:
        if s[i] == ".":
            gaps += 1
            nt = 0
        else:
            if gaps >= THRESH:
                total = len(s) - i
            gaps = 0
            nt += 1
        i -= 1
    return total


def gap_stretch_keys(C, R):
    """Streaming form of findGapStretches: returns the final dict as a list of (start, end)
    in insertion order of distinct keys (values overwritten like the dict)."""
    L = len(C)
    nslots = 0            # len(positionsStretch), empty slots included
    cur = None            # last slot: None = empty / absent, else [a, b]
    have_cur = False      # a last slot exists
    cg = cr = 0
    prev = None
    d = {}
    state = {"pend": None, "merge": False, "k": 0}

    def emit2(a, b):
        if a == 0:
            if b - a > THRESH2:
                d[0] = b
        elif b == L - 1:
            if b - a > THRESH2:
                d[a] = b

    def emit_tmp(a, b):
        p = state["pend"]
        if p is not None:
            if a - p[1] <= THRESH:
                emit2(p[0], b)
                state["merge"] = True
            else:
                emit2(p[0], p[1])
                state["merge"] = False
        state["pend"] = (a, b)
        state["k"] += 1

    def finalize(slot, many):
        if slot is None:
            return
        a, b = slot
        if many:
            if a <= THRESH2:
                emit_tmp(0, b)
            if L - b <= THRESH2:
                emit_tmp(a, L - 1)
            else:
                emit_tmp(a, b)
        else:
            e = [0, b] if a <= THRESH2 else [a, b]
            if L - b <= THRESH2:
                e[1] = L - 1
            emit_tmp(e[0], e[1])

    for pos in range(L):
        c, r = C[pos], R[pos]
        if prev == ".":
            if c == ".":
                cg = cg + 1 if cg > 0 else 2
            if r == ".":
                cr = cr + 1 if cr > 0 else 2
        if prev is None:
            if c == ".":
                cg += 1
            if r == ".":
                cr += 1
        if c != ".":
            if cg > 0:
                if have_cur:
                    finalize(cur, True)   # a second slot exists from now on
                cur, have_cur = None, True
                nslots += 1
            cg = 0
        if r != ".":
            cr = 0
        if cg >= THRESH and cr < THRESH2:
            if nslots == 0:
                cur, have_cur, nslots = [pos - THRESH + 1, pos], True, 1
            else:
                if cur is None:
                    cur = [pos - THRESH + 1, pos]
                cur[1] = pos
        prev = c
    if have_cur:
        finalize(cur, nslots > 1)
    if state["pend"] is not None and not state["merge"]:
        emit2(*state["pend"])
    return list(d.items())


def tally_read(R, C, U):
    """returns dict over FIELDS for one merged read (rows of equal length)"""
    L = len(R)
    out = dict.fromkeys(FIELDS, 0)
    out["ncols"] = L
    out["extended"] = -1
    if L <= 10:
        return out
    out["assessed"] = 1
    gl = min(nb_left_gaps(R), nb_left_gaps(U))
    gr = min(nb_right_gaps(R), nb_right_gaps(U))
    mask = [True] * L
    ext = -1
    if gl >= THRESH:
        for i in range(gl):
            mask[i] = False
        if gl >= THRESH2:
            ext = max(ext, 0) + gl - C[:gl].count(".")
    if gr >= THRESH:
        for i in range(L - 1, L - gr, -1):
            mask[i] = False
        if gr >= THRESH2:
            ext = max(ext, 0) + gr - C[L - gr + 1:].count(".")
    keys = gap_stretch_keys(C, R)
    missing = 0
    for a, b in keys:
        missing += b - a - R[a:b + 1].count(".")
        for i in range(a, b + 1):
            mask[i] = False
    missing = max(0, missing - (gl + gr))
    for i in range(L):
        r, c, u = R[i], C[i], U[i]
        if r in "gcGC":
            out["GCref"] += 1
        if c in "gcGC":
            out["GCcor"] += 1
        if not mask[i]:
            continue
        if c != r:
            if r == ".":
                out["insC"] += 1
            elif c != ".":
                out["subsC"] += 1
            else:
                out["delC"] += 1
        if u != r:
            if r == ".":
                out["insU"] += 1
            elif u != ".":
                out["subsU"] += 1
            else:
                out["delU"] += 1
        if r == u:
            if u != c:
                out["FP"] += 1; out["uncor"] += 1
            else:
                out["TP"] += 1; out["cor"] += 1
            out["uncorCor"] += 1
        else:
            if r == c:
                out["TP"] += 1; out["cor"] += 1
            else:
                if u == c:
                    out["FN"] += 1; out["FP"] += 1
                out["uncor"] += 1
            out["uncorUncor"] += 1
    out["lenRef"] = L - R.count(".")
    out["lenCor"] = L - C.count(".")
    out["lenUnc"] = L - U.count(".")
    out["gapsLeft"], out["gapsRight"], out["missing"], out["extended"] = gl, gr, missing, ext
    return out


def merge_windows(records):
    """records: list of (headers[3], rows[3]) in PIR order (ref, corrected, uncorrected).
    Returns list of (header, R, C, U) exactly as Donatello writes them (header already cut)."""
    out = []
    if not records:
        return out

    def clean(a, b, c):
        keep = [i for i, ch in enumerate(b) if ch != "n"]
        return ("".join(a[i] for i in keep if i < len(a)), "".join(b[i] for i in keep),
                "".join(c[i] for i in keep if i < len(c)))

    header = records[0][0][2]
    acc = [records[0][1][0], records[0][1][1], records[0][1][2]]
    # the reference loops once more after the last record with an empty header line
    for heads, rows in records[1:] + [(("", "", ""), records[-1][1])]:
        if header != heads[2]:
            if len(acc[0]) > 1:
                a, b, c = clean(*acc)
                out.append((header[:len(header) - 11] + " ", a, b, c))
                header = heads[2]
            acc = [rows[0], rows[1], rows[2]]
        else:
            acc = [acc[0] + rows[0], acc[1] + rows[1], acc[2] + rows[2]]
    return out

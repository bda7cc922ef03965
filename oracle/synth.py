"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic inputs for the POA path.

Two generators:
  * hard_windows(): adversarial window triplets (tiny alphabets => many score ties,
    high error rates, trimmed ends, `N` / `AAA` placeholders as the reference
    splitter emits them (src/split/Master_Splitter.cpp:139-154,417-423), IUPAC and
    other out-of-alphabet letters, mixed case, lengths 1..~500).
  * reads(): read-level triplets for the BASELINE.json configs:
    reference = i.i.d. uniform ACGT, raw = reference with error events, corrected =
    the same events each kept with probability p_keep.
PRNG: splitmix64, so the streams are reproducible in C as well.
"""
import math

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def below(self, n):
        return self.next() % n

    def unit(self):
        return (self.next() >> 11) / float(1 << 53)


def mutate(rng, seq, rate, alphabet, ins=1, dele=1, sub=1):
    out = []
    tot = ins + dele + sub
    for ch in seq:
        if rng.unit() < rate:
            k = rng.below(tot)
            if k < ins:
                out.append(alphabet[rng.below(len(alphabet))])
                out.append(ch)
            elif k < ins + dele:
                pass
            else:
                out.append(alphabet[rng.below(len(alphabet))])
        else:
            out.append(ch)
    return "".join(out)


def hard_windows(n, seed=7):
    """returns list of (name, ref, cor, unc)"""
    rng = SplitMix64(seed)
    wins = []
    alphabets = ["ACGT", "AC", "A", "ACGTN", "acgt", "ACG"]
    for i in range(n):
        ab = alphabets[rng.below(len(alphabets))]
        mode = rng.below(12)
        if mode == 0:
            L = 1 + rng.below(6)
        elif mode == 1:
            L = 100 + rng.below(400)
        else:
            L = 8 + rng.below(90)
        ref = "".join(ab[rng.below(len(ab))] for _ in range(L))
        er_u = [0.01, 0.1, 0.2, 0.3][rng.below(4)]
        er_c = [0.0, 0.01, 0.05, 0.3][rng.below(4)]
        unc = mutate(rng, ref, er_u, ab) or ab[0]
        cor = mutate(rng, ref, er_c, ab) or ab[0]
        t = rng.below(16)
        if t == 0:
            cor = "N"
        elif t == 1:
            cor = cor[: max(1, len(cor) // 3)]
        elif t == 2:
            cor = cor[len(cor) // 2:] or "N"
        elif t == 3:
            ref, cor, unc = "AAA", "AAA", "AAA"
        elif t == 4:
            unc = unc[: max(1, len(unc) // 4)]
        elif t == 5:  # out-of-alphabet letters and mixed case
            extra = "RYKMSWBDHVN?]-x"
            l = list(cor)
            for _ in range(1 + rng.below(4)):
                l[rng.below(len(l))] = extra[rng.below(len(extra))]
            cor = "".join(l)
            unc = "".join(c.lower() if rng.below(2) else c for c in unc)
        elif t == 6:  # unrelated sequences
            cor = "".join(ab[rng.below(len(ab))] for _ in range(1 + rng.below(60)))
        elif t == 7:
            unc = "".join(ab[rng.below(len(ab))] for _ in range(1 + rng.below(60)))
        wins.append(("w%d" % i, ref, cor, unc))
    return wins


def write_windows(wins, prefix, titles=False):
    """writes prefix.ref.fa / .cor.fa / .unc.fa in the shard format of the splitter"""
    with open(prefix + ".ref.fa", "w") as fr, open(prefix + ".cor.fa", "w") as fc, open(prefix + ".unc.fa", "w") as fu:
        for k, (name, r, c, u) in enumerate(wins):
            h = ">" + name + (" some title %d" % k if titles and k % 3 == 0 else "")
            fr.write(h + "\n" + r + "\n")
            fc.write(h + "\n" + c + "\n")
            fu.write(h + "\n" + u + "\n")


def reads(n, length_fn, seed, raw_rate=0.10, keep=0.1, ratios=(1, 1, 1), homopolymer_del_bias=False):
    """read-level triplets: yields (name, ref, raw, corrected)."""
    rng = SplitMix64(seed)
    ab = "ACGT"
    ins, dele, sub = ratios
    tot = ins + dele + sub
    for i in range(n):
        L = length_fn(rng)
        ref = [ab[rng.next() & 3] for _ in range(L)]
        raw = []
        cor = []
        prev = ""
        for ch in ref:
            rate = raw_rate
            if homopolymer_del_bias and ch == prev:
                rate = min(0.9, raw_rate * 1.5)
            prev = ch
            if rng.unit() < rate:
                k = rng.below(tot)
                kept = rng.unit() < keep
                if k < ins:
                    b = ab[rng.next() & 3]
                    raw.append(b); raw.append(ch)
                    if kept:
                        cor.append(b)
                    cor.append(ch)
                elif k < ins + dele:
                    if not kept:
                        cor.append(ch)
                else:
                    b = ab[rng.next() & 3]
                    raw.append(b)
                    cor.append(b if kept else ch)
            else:
                raw.append(ch)
                cor.append(ch)
        yield ("read_%d" % i, "".join(ref), "".join(raw), "".join(cor))


def loguniform(lo, hi):
    return lambda rng: int(math.exp(math.log(lo) + rng.unit() * (math.log(hi) - math.log(lo))))


def random_msa_rows(n, seed=21):
    """gap-rich random 3-row MSAs (ref, corrected, uncorrected) for the tally: long '.' runs at the
    borders (trimmed / extended reads), gap stretches inside, 'n' never present (Donatello drops
    those columns).  Returns list of (R, C, U) strings of equal length."""
    rng = SplitMix64(seed)
    out = []
    for _ in range(n):
        L = 11 + rng.below(400) if rng.below(8) else 1 + rng.below(14)
        rows = []
        base = [("acgt")[rng.below(4)] for _ in range(L)]
        for k in range(3):
            row = []
            p_gap = [0.02, 0.05, 0.15][rng.below(3)]
            p_mut = [0.0, 0.02, 0.2][rng.below(3)]
            i = 0
            while i < L:
                if rng.unit() < p_gap / 4:
                    run = 1 + rng.below([3, 8, 30, 60][rng.below(4)])
                    row.extend("." * min(run, L - i))
                    i += min(run, L - i)
                else:
                    row.append("acgt"[rng.below(4)] if rng.unit() < p_mut else base[i])
                    i += 1
            # border effects
            t = rng.below(10)
            if t == 0:
                g = min(L, 3 + rng.below(40)); row[:g] = "." * g
            elif t == 1:
                g = min(L, 3 + rng.below(40)); row[L - g:] = "." * g
            elif t == 2:
                g = min(L // 2, 3 + rng.below(30)); row[:g] = "." * g; row[L - g:] = "." * g
            row = row[:L]
            if all(ch == "." for ch in row):  # the reference divides by the non-gap length
                row[rng.below(L)] = base[0]
            rows.append("".join(row))
        out.append((rows[0], rows[1], rows[2]))
    return out

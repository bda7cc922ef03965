"""CPU: the per-window DEVICE code (elector_b200/csrc/poa_kernel.cuh) compiled as host code
and run by tests/emul/poa_emul.cu, against the reference goldens.  This checks the kernel's
restructuring (8-row register bands, two frontier sets, 2-bit moves + ordinals, fused
emit) without a GPU; the GPU tests then check the real launch path."""
import subprocess

import pytest

from conftest import GOLDEN_SETS, parse_dump


@pytest.mark.parametrize("mode", ["int32", "packed", "dual-general", "coop"])
@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_emulated_kernel_equals_reference(golden_dir, emul_bin, name, mode, tmp_path):
    """mode packed: the 16-bit packed kernels wherever the library would use them (Phase1P, Phase2L for windows whose P1 is
    linear, the dual-frontier Phase2D of poa_dual.cuh for the rest); dual-general: Phase2D for every window; coop: both DPs
    of every window through the warp-cooperative wavefront of poa_coop.cuh (32 lanes of 8 rows, emulated lane by lane).
    The emulator also checks that the number of MSA columns announced before the fusion equals the number emitted."""
    d = golden_dir
    pir, sc = str(tmp_path / "e.pir"), str(tmp_path / "e.scores")
    cmd = [emul_bin, d + "/blosum80.mat", "%s/%s.ref.fa" % (d, name), "%s/%s.cor.fa" % (d, name),
           "%s/%s.unc.fa" % (d, name), pir, sc] + ([mode] if mode != "int32" else [])
    assert subprocess.call(cmd) == 0
    assert open(pir, "rb").read() == open("%s/%s.pir" % (d, name), "rb").read()
    gold = parse_dump("%s/%s.dump" % (d, name))
    got = [tuple(int(v) for v in line.split()) for line in open(sc)]
    assert len(got) == len(gold)
    for g, e in zip(gold, got):
        assert (g["s1"], g["s2"], g["n1"]) == e[:3]


@pytest.mark.parametrize("band_w", ["1", "3", "6", "16"])
@pytest.mark.parametrize("name", ["hard", "example_tail"])
def test_emulated_band_dp_is_exact(golden_dir, emul_bin, name, band_w, tmp_path, monkeypatch):
    """the diagonal-band DP of the packed linear kernels with its exactness test (poa_packed.cuh): whatever the half-width,
    a window either passes the test and equals the reference, or is run again without the band"""
    monkeypatch.setenv("ELECTOR_BAND_W", band_w)
    d = golden_dir
    pir, sc = str(tmp_path / "e.pir"), str(tmp_path / "e.scores")
    cmd = [emul_bin, d + "/blosum80.mat", "%s/%s.ref.fa" % (d, name), "%s/%s.cor.fa" % (d, name), "%s/%s.unc.fa" % (d, name), pir, sc, "packed"]
    assert subprocess.call(cmd) == 0
    assert open(pir, "rb").read() == open("%s/%s.pir" % (d, name), "rb").read()
    gold = parse_dump("%s/%s.dump" % (d, name))
    got = [tuple(int(v) for v in line.split()) for line in open(sc)]
    assert [(g["s1"], g["s2"], g["n1"]) for g in gold] == [e[:3] for e in got]


def test_emulated_kernel_generic_matrix(golden_dir, emul_bin, tmp_path):
    """non-uniform substitution scores + other gap penalties: table path vs the oracle"""
    from elector_b200.matrix import ALPHABET
    from oracle import oracle
    d = golden_dir
    mp = str(tmp_path / "m.mat")
    lines = ["GAP-TRUNCATION-LENGTH=4", "GAP-DECAY-LENGTH=0", "GAP-PENALTIES=7 3 3", "  " + " ".join(ALPHABET)]
    for i, a in enumerate(ALPHABET):
        lines.append(a + " " + " ".join(str(4 if i == j else -((i * 7 + j * 3) % 5) - 1) for j in range(len(ALPHABET))))
    open(mp, "w").write("\n".join(lines) + "\n")
    pir, opir = str(tmp_path / "e.pir"), str(tmp_path / "o.pir")
    assert subprocess.call([emul_bin, mp, d + "/hard.ref.fa", d + "/hard.cor.fa", d + "/hard.unc.fa", pir]) == 0
    assert oracle.poa_files(mp, d + "/hard.ref.fa", d + "/hard.cor.fa", d + "/hard.unc.fa", opir) == 0
    assert open(pir, "rb").read() == open(opir, "rb").read()


def test_emulated_kernel_rejects_unsupported_matrix(golden_dir, emul_bin, tmp_path):
    from elector_b200.matrix import default_matrix_text
    mp = str(tmp_path / "m.mat")
    open(mp, "w").write(default_matrix_text(gaps=(10, 5, 1)))  # decaying extension penalty
    d = golden_dir
    rc = subprocess.call([emul_bin, mp, d + "/edge.ref.fa", d + "/edge.cor.fa", d + "/edge.unc.fa", str(tmp_path / "x.pir")],
                         stderr=subprocess.DEVNULL)
    assert rc == 3


@pytest.mark.parametrize("mode", ["int32", "packed", "dual-general", "coop"])
def test_emulated_kernel_long_windows(golden_dir, emul_bin, mode, tmp_path):
    """many bands, 8- and 16-row last bands, long placeholder / trimmed windows, vs the oracle"""
    from oracle import oracle, synth
    rng = synth.SplitMix64(4242)
    wins = []
    for L in (15, 16, 17, 24, 25, 31, 32, 33, 40, 41, 129, 257, 300, 511, 700, 1100):
        ref = "".join("ACGT"[rng.below(4)] for _ in range(L))
        wins.append((ref, synth.mutate(rng, ref, 0.03, "ACGT") or "A", synth.mutate(rng, ref, 0.12, "ACGT") or "A"))
        wins.append((ref, synth.mutate(rng, ref, 0.3, "AC") or "A", synth.mutate(rng, ref, 0.3, "AC") or "A"))
    wins.append((wins[-2][0], "N", wins[-2][2]))
    wins.append((wins[-3][0], wins[-3][1][:40], wins[-3][2]))
    wins.append(("ACGT" * 30, "ACGT" * 30, "A"))
    paths = []
    for k, key in enumerate(("ref", "cor", "unc")):
        p = str(tmp_path / (key + ".fa"))
        with open(p, "w") as f:
            for i, w in enumerate(wins):
                f.write(">w%d\n%s\n" % (i, w[k]))
        paths.append(p)
    pir, opir = str(tmp_path / "e.pir"), str(tmp_path / "o.pir")
    mp = golden_dir + "/blosum80.mat"
    assert subprocess.call([emul_bin, mp, paths[0], paths[1], paths[2], pir] + (["-", mode] if mode != "int32" else [])) == 0
    assert oracle.poa_files(mp, paths[0], paths[1], paths[2], opir) == 0
    assert open(pir, "rb").read() == open(opir, "rb").read()


@pytest.mark.parametrize("mode", ["packed", "dual-general"])
def test_emulated_dual_kernel_on_bubble_windows(golden_dir, emul_bin, mode, tmp_path):
    """the skewed two-set dual kernel (poa_dual.cuh) on ELECTOR-like windows: a corrected sequence with a few differences (bubbles
    that open and close at every position of a band, in the low and in the high half, back to back, at the first and at the last
    node; late starts and early ends of either sequence) against a noisier uncorrected one, lengths around the band sizes"""
    from oracle import oracle, synth
    rng = synth.SplitMix64(20261018)
    wins = []
    for i in range(6000):
        L = [7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 32, 33, 40, 48, 50, 55, 64, 65, 80, 100, 127][rng.below(21)]
        ref = "".join("ACGT"[rng.below(4)] for _ in range(L))
        cor = synth.mutate(rng, ref, [0.01, 0.02, 0.05, 0.1, 0.2][rng.below(5)], "ACGT")
        t = rng.below(8)
        if t == 0:
            cor = cor[1 + rng.below(3):]                      # cor starts late: INITIAL node with the virtual link first
        elif t == 1:
            cor = cor[:max(1, len(cor) - 1 - rng.below(3))]   # cor ends early: several FINAL nodes
        elif t == 2:
            cor = "ACGT"[rng.below(4)] * (1 + rng.below(3)) + cor   # cor starts with an insertion
        unc = synth.mutate(rng, ref, [0.05, 0.1, 0.2][rng.below(3)], "ACGT")
        wins.append((ref, cor or "A", unc or "C"))
    paths = []
    for k, key in enumerate(("ref", "cor", "unc")):
        p = str(tmp_path / (key + ".fa"))
        with open(p, "w") as f:
            for i, w in enumerate(wins):
                f.write(">w%d\n%s\n" % (i, w[k]))
        paths.append(p)
    pir, opir = str(tmp_path / "e.pir"), str(tmp_path / "o.pir")
    mp = golden_dir + "/blosum80.mat"
    assert subprocess.call([emul_bin, mp, paths[0], paths[1], paths[2], pir, "-", mode], stderr=subprocess.DEVNULL) == 0
    assert oracle.poa_files(mp, paths[0], paths[1], paths[2], opir) == 0
    assert open(pir, "rb").read() == open(opir, "rb").read()

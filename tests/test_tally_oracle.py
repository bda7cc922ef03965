"""CPU: the merge + tally restatement (oracle/tally_oracle.py) against outputs of the reference's
own computeStats.py functions and of the compiled Donatello (tests/golden/tally_*, donatello.*)."""
import gzip
import json
import os

from conftest import GOLD, parse_pir

INT_FIELDS = ["TP", "FP", "FN", "cor", "uncor", "uncorCor", "uncorUncor", "insC", "delC", "subsC", "insU", "delU",
              "subsU", "lenRef", "lenCor", "lenUnc", "gapsLeft", "gapsRight", "missing", "extended", "ncols", "assessed"]


def load(name):
    return json.loads(gzip.open(os.path.join(GOLD, name)).read())


def check(got, exp):
    if not exp["assessed"]:
        assert got["assessed"] == 0 and got["ncols"] == exp["ncols"]
        return
    for k in INT_FIELDS:
        assert got[k] == exp[k], (k, got[k], exp[k])
    assert round(got["GCref"] * 1.0 / got["lenRef"], 3) == exp["GCrateRef"]
    assert round(got["GCcor"] * 1.0 / got["lenCor"], 3) == exp["GCrateCor"]


def test_tally_oracle_on_example_reads():
    from oracle import tally_oracle as to
    ex = load("tally_example.json.gz")
    assert len(ex) >= 30
    for e in ex:
        check(to.tally_read(e["R"], e["C"], e["U"]), e["expect"])
        if e["expect"]["assessed"]:
            assert sorted([a, b] for a, b in to.gap_stretch_keys(e["C"], e["R"])) == e["expect"]["stretches"]


def test_tally_oracle_on_random_gap_rich_rows():
    from oracle import synth, tally_oracle as to
    exp = load("tally_random.json.gz")
    rows = synth.random_msa_rows(3000, 21)
    assert len(rows) == len(exp)
    for (R, C, U), e in zip(rows, exp):
        check(to.tally_read(R, C, U), e)
        if e["assessed"]:
            assert sorted([a, b] for a, b in to.gap_stretch_keys(C, R)) == e["stretches"]


def test_merge_oracle_equals_donatello(golden_dir):
    from oracle import tally_oracle as to
    recs = parse_pir(golden_dir + "/donatello.pir")
    merged = to.merge_windows(recs)
    exp = open(golden_dir + "/donatello.msa").read().split("\n")
    got = []
    for h, a, b, c in merged:
        got += [h, a, h, b, h, c]
    assert got == exp[:-1]

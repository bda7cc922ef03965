"""CPU: the WHOLE README example (484 triplets, 82 804 windows, 5.6e8 DP cells) through the oracle chain against what the
unmodified reference made of it (tests/golden/example_full.json.gz, oracle/make_golden_example.py): every shard's PIR,
the merged msa.fa, the per-read counters of all 484 merged records."""
import hashlib
import os
import subprocess

from conftest import ROOT, md5_file, parse_pir


def test_sorted_inputs_and_reference_splitter_reproduce_the_golden(example_chain):
    ex = example_chain
    g = ex["gold"]
    for k in ("ref", "unc", "cor"):
        assert md5_file(os.path.join(ex["work"], k + ".fa")) == g["sorted"][k]
    assert ex["shards"] == g["shards"]
    for i in ex["shards"]:
        for k in ("out1", "out2", "out3"):
            assert md5_file("%s/%s%d" % (ex["out"], k, i)) == g["splitter"][str(i)][k]
    assert int(open(ex["out"] + "/small_reads.txt").read()) == g["small_reads"]


def test_oracle_chain_equals_reference_on_every_example_window(example_chain, example_oracle_msa):
    from oracle import tally_oracle as to
    ex = example_chain
    g = ex["gold"]
    # poa: byte-identical PIR of every shard
    for i in ex["shards"]:
        assert md5_file("%s/smsa%d" % (example_oracle_msa["dir"], i)) == g["smsa_md5"][str(i)], i
    # Donatello: byte-identical msa.fa (shard by shard, appended, alignment.py:121-127)
    recs = example_oracle_msa["records"]
    msa = []
    for h, a, b, c in recs:
        msa += [h, a, h, b, h, c]
    assert hashlib.md5(("\n".join(msa) + "\n").encode()).hexdigest() == g["msa_md5"]
    # computeStats: the integer counters of ALL merged records
    assert len(recs) == len(g["records"]) == 484
    n_assessed = 0
    for (h, R, C, U), e in zip(recs, g["records"]):
        assert h == e["header"]
        got, exp = to.tally_read(R, C, U), e["expect"]
        assert got["ncols"] == exp["ncols"] and got["assessed"] == exp["assessed"]
        if not exp["assessed"]:
            continue
        n_assessed += 1
        for k in ("TP", "FP", "FN", "cor", "uncor", "uncorCor", "uncorUncor", "insC", "delC", "subsC", "insU", "delU", "subsU",
                  "lenRef", "lenCor", "lenUnc", "gapsLeft", "gapsRight", "missing", "extended"):
            assert got[k] == exp[k], (h, k, got[k], exp[k])
        assert sorted([a, b] for a, b in to.gap_stretch_keys(C, R)) == exp["stretches"]
    assert n_assessed == 459

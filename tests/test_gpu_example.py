"""GPU: the WHOLE README example through the CUDA path against what the unmodified reference made of it
(tests/golden/example_full.json.gz): the `poa` drop-in on every shard (PIR md5), then one pipelined call over all
82 804 windows -> msa.fa bytes (md5) and the integer counters of all 484 merged records (459 assessed)."""
import hashlib
import subprocess

import numpy as np
import pytest

from conftest import md5_file, read_fasta_simple

pytestmark = pytest.mark.gpu
INT_FIELDS = ["TP", "FP", "FN", "cor", "uncor", "uncorCor", "uncorUncor", "insC", "delC", "subsC", "insU", "delU",
              "subsU", "lenRef", "lenCor", "lenUnc", "gapsLeft", "gapsRight", "missing", "extended", "ncols", "assessed"]


def test_poa_dropin_pir_md5_of_every_example_shard(example_chain, golden_dir, tmp_path):
    """alignment.py:60's command line on each of the 10 shards; bytes == the reference poa's (md5 in the golden)"""
    from elector_b200.lib import poa_binary_path
    ex = example_chain
    for i in ex["shards"]:
        out = str(tmp_path / ("smsa%d" % i))
        p = subprocess.run([poa_binary_path(), "-pir", out, "-preserve_seqorder", "-corrected_reads_fasta", "%s/out3%d" % (ex["out"], i),
                            "-reference_reads_fasta", "%s/out1%d" % (ex["out"], i), "-uncorrected_reads_fasta", "%s/out2%d" % (ex["out"], i),
                            "-preserve_seqorder", "-threads", "1", "-pathMatrix", golden_dir + "/blosum80.mat"], capture_output=True)
        assert p.returncode == 0 and p.stderr == b""
        assert md5_file(out) == ex["gold"]["smsa_md5"][str(i)], i


def test_pipeline_msa_md5_and_all_459_read_counters(example_chain):
    import elector_b200
    from elector_b200 import TALLY_FIELDS, windows_to_csr
    ex = example_chain
    g = ex["gold"]
    heads, refs, cors, uncs, shard_first = [], [], [], [], [0]
    for i in ex["shards"]:
        r = read_fasta_simple("%s/out1%d" % (ex["out"], i))
        c = read_fasta_simple("%s/out3%d" % (ex["out"], i))
        u = read_fasta_simple("%s/out2%d" % (ex["out"], i))
        assert len(r) == len(c) == len(u) == g["splitter_records"][str(i)]
        heads += [h for h, _ in r]; refs += [s for _, s in r]; cors += [s for _, s in c]; uncs += [s for _, s in u]
        shard_first.append(len(heads))
    assert len(heads) == 82804
    # reads = runs of equal headers, never across a shard file (Donatello runs per smsa file)
    bounds = set(shard_first)
    first = [0] + [k for k in range(1, len(heads)) if heads[k] != heads[k - 1] or k in bounds] + [len(heads)]
    first = sorted(set(first))
    (r, ro), (c, co), (u, uo) = windows_to_csr(refs), windows_to_csr(cors), windows_to_csr(uncs)
    with elector_b200.PoaContext(0) as ctx:
        res, counters, sums = ctx.pipeline_csr(r, ro, c, co, u, uo, first)
        merged = ctx.merge(res, first)
    # Donatello's text: header = PIR header (">name untitled") minus its last 11 characters, plus a blank
    msa = []
    kept = []
    for k, (a, b, cc) in enumerate(merged):
        if len(a) <= 1:        # Donatello.cpp:70 writes nothing for such a read
            continue
        h = ">" + heads[first[k]] + " untitled"
        h = h[:len(h) - 11] + " "
        msa += [h, a, h, b, h, cc]
        kept.append(k)
    assert hashlib.md5(("\n".join(msa) + "\n").encode()).hexdigest() == g["msa_md5"]
    assert len(kept) == len(g["records"]) == 484
    n_assessed = 0
    for k, e in zip(kept, g["records"]):
        got = dict(zip(TALLY_FIELDS, (int(v) for v in counters[k])))
        exp = e["expect"]
        assert got["ncols"] == exp["ncols"] and got["assessed"] == exp["assessed"]
        if not exp["assessed"]:
            continue
        n_assessed += 1
        for f in INT_FIELDS:
            assert got[f] == exp[f], (e["header"], f, got[f], exp[f])
        assert round(got["GCref"] * 1.0 / got["lenRef"], 3) == exp["GCrateRef"]
        assert round(got["GCcor"] * 1.0 / got["lenCor"], 3) == exp["GCrateCor"]
    assert n_assessed == 459
    ext = TALLY_FIELDS.index("extended")
    exp_sums = counters.sum(axis=0)
    exp_sums[ext] = counters[:, ext][counters[:, ext] >= 0].sum()
    assert np.array_equal(sums, exp_sums)


def test_example_from_the_read_files_to_the_report(example_chain, tmp_path):
    """SURVEY.md 8f-1 + 8f-2 in one go: the three sorted read files -> elector_reads_run (window cutting, alignment, Donatello merge and
    tally, windows never leave the device) -> elector_report_write.  msa.fa bytes, per_read_metrics.txt (md5 5507a652...),
    read_size_distribution.txt, the log block and the printed block == what the unmodified reference chain wrote."""
    import os

    import elector_b200
    import workloads
    from test_report import README_BLOCK, check_against_golden
    ex = example_chain
    g, work = ex["gold"], ex["work"]
    hr, ref, ro = workloads.parse_two_line_fasta(work + "/ref.fa")
    _, unc, uo = workloads.parse_two_line_fasta(work + "/unc.fa")
    _, cor, co = workloads.parse_two_line_fasta(work + "/cor.fa")
    n = len(hr)
    assert n == 484
    with elector_b200.PoaContext(0) as ctx:
        got = ctx.reads_run(ref, ro, unc, uo, cor, co, [len(h) for h in hr], 0.1, merged=True)
        stretches = ctx.last_stretches(n)
        assert got["n_windows"] == 82804
        small, wrong = int((got["status"] == 1).sum()), int((got["status"] == 2).sum())
        assert (small, wrong) == (g["small_reads"], g["wrongly_cor_reads"])
        heads, rows, keep = [], [], []
        for t in range(n):
            o, l = int(got["m_off"][t]), int(got["m_len"][t])
            if l <= 1:                      # Donatello.cpp:70 writes nothing for such a read
                continue
            h = hr[t] if isinstance(hr[t], str) else hr[t].decode()
            h = (h if h.startswith(">") else ">" + h) + " untitled"
            heads.append(h[:len(h) - 11] + " ")
            rows.append(tuple(got[k][o:o + l].tobytes().decode() for k in ("m_ref", "m_cor", "m_unc")))
            keep.append(t)
        msa = []
        for h, (a, b, c) in zip(heads, rows):
            msa += [h, a, h, b, h, c]
        assert hashlib.md5(("\n".join(msa) + "\n").encode()).hexdigest() == g["msa_md5"]
        out1, out2 = str(tmp_path / "a"), str(tmp_path / "b")
        os.makedirs(out1); os.makedirs(out2)
        common = dict(small_reads=small, wrongly_cor_reads=wrong, size_threshold=0.1, homopolymer_threshold=5, corrected_fasta=work + "/cor.fa")
        # from the counters of the chained call
        res = elector_b200.report_write([h[1:] for h in heads], got["counters"][keep], stretches[keep], [r[0] for r in rows], [r[1] for r in rows],
                                        out_dir=out1, compensated_sum=True, **common)
        check_against_golden(res, out1, g)
        # from the rows alone (msa.fa in memory): tally on the device, then the report
        res = elector_b200.report_run(ctx, [h[1:] for h in heads], [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows],
                                      out_dir=out2, compensated_sum=False, **common)
        assert res["stdout"] == "None\n" + README_BLOCK
        assert md5_file(out2 + "/per_read_metrics.txt") == g["report"]["per_read_metrics_md5"]

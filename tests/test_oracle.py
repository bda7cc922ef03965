"""CPU: the oracle (oracle/poa_oracle.c) against outputs of the compiled reference."""
import hashlib
import os
import subprocess

import pytest

from conftest import GOLDEN_SETS, ROOT


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_oracle_pir_equals_reference(golden_dir, name, tmp_path):
    from oracle import oracle
    d = golden_dir
    out = str(tmp_path / "o.pir")
    rc = oracle.poa_files(d + "/blosum80.mat", "%s/%s.ref.fa" % (d, name), "%s/%s.cor.fa" % (d, name),
                          "%s/%s.unc.fa" % (d, name), out)
    assert rc == 0
    assert open(out, "rb").read() == open("%s/%s.pir" % (d, name), "rb").read()


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_oracle_scores_and_maps_equal_reference(golden_dir, name, tmp_path):
    """both align_lpo_po scores and the four x_to_y / y_to_x maps (align_lpo_po2.c:486,158-165)"""
    from oracle import oracle
    d = golden_dir
    out = str(tmp_path / "o.dump")
    assert oracle.dump_files(d + "/blosum80.mat", "%s/%s.ref.fa" % (d, name), "%s/%s.cor.fa" % (d, name),
                             "%s/%s.unc.fa" % (d, name), out) == 0
    assert open(out).read() == open("%s/%s.dump" % (d, name)).read()


def test_oracle_batch_matches_file_driver(golden_dir):
    from conftest import parse_pir, read_fasta_simple
    from elector_b200 import windows_to_csr
    from oracle import oracle
    d = golden_dir
    refs = [s for _, s in read_fasta_simple(d + "/hard.ref.fa")]
    cors = [s for _, s in read_fasta_simple(d + "/hard.cor.fa")]
    uncs = [s for _, s in read_fasta_simple(d + "/hard.unc.fa")]
    r, ro = windows_to_csr(refs); c, co = windows_to_csr(cors); u, uo = windows_to_csr(uncs)
    o = oracle.batch(r, ro, c, co, u, uo, nthreads=4)
    gold = parse_pir(d + "/hard.pir")
    for w in range(len(refs)):
        assert oracle.window_rows(o, w) == gold[w][1]


def test_oracle_cli_flags(golden_dir, tmp_path):
    """oracle CLI has the reference's five flags and ignores the others (main.c:85-113)"""
    from oracle import oracle
    oracle.build()
    d = golden_dir
    out = str(tmp_path / "cli.pir")
    rc = subprocess.call([os.path.join(ROOT, "oracle", "poa_oracle_cli"), "-pir", out, "-preserve_seqorder",
                          "-corrected_reads_fasta", d + "/edge.cor.fa", "-reference_reads_fasta", d + "/edge.ref.fa",
                          "-uncorrected_reads_fasta", d + "/edge.unc.fa", "-preserve_seqorder", "-threads", "1",
                          "-pathMatrix", d + "/blosum80.mat"], stdout=subprocess.DEVNULL)
    assert rc == 0
    assert open(out, "rb").read() == open(d + "/edge.pir", "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "poa")), reason="compiled reference not present")
def test_live_reference_agrees_with_golden_and_generated_matrix(golden_dir, tmp_path):
    """when oracle/_ref/poa exists: it reproduces the committed golden with OUR generated matrix file"""
    d = golden_dir
    out = str(tmp_path / "ref.pir")
    rc = subprocess.call([os.path.join(ROOT, "oracle", "_ref", "poa"), "-pir", out, "-corrected_reads_fasta", d + "/hard.cor.fa",
                          "-reference_reads_fasta", d + "/hard.ref.fa", "-uncorrected_reads_fasta", d + "/hard.unc.fa",
                          "-pathMatrix", d + "/blosum80.mat"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert rc == 0
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == hashlib.md5(open(d + "/hard.pir", "rb").read()).hexdigest()


def test_oracle_long_windows_equal_reference(golden_dir, tmp_path):
    """golden set `long` (oracle/make_golden_long.py): a 3 300-letter window, a 33 500-letter window on one FASTA line
    longer than the reference's 32 KiB line buffer (fasta_format.c:20), wrapped lines with blanks, a window after them"""
    from conftest import parse_dump
    from oracle import oracle
    d = golden_dir
    out, dump = str(tmp_path / "o.pir"), str(tmp_path / "o.dump")
    assert oracle.poa_files(d + "/blosum80.mat", d + "/long.ref.fa", d + "/long.cor.fa", d + "/long.unc.fa", out) == 0
    assert open(out, "rb").read() == open(d + "/long.pir", "rb").read()

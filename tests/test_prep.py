"""CPU: readAndSortFasta / duplicateRefReads (elector/readAndSortFiles.py:150-191) through the C-ABI (host code):
  - the README example: the three files it writes == the md5s of the golden's prepared files (tests/golden/example_full.json.gz,
    "sorted": the inputs from which the unmodified splitter / poa / Donatello / computeStats chain reproduced README.md:137-161);
  - FASTA shapes Bio.SeqIO accepts (multi-line and blank-padded sequences, CRLF, text before the first header, equal headers,
    reads without a corrected read) against a restatement of SimpleFastaParser + the reference's two functions in this file."""
import os
import re

from conftest import load_example_golden, md5_file


def simple_fasta(path):
    """Bio.SeqIO.FastaIO.SimpleFastaParser"""
    recs, title, lines = [], None, []
    for line in open(path, newline=""):
        if line[0:1] == ">":
            if title is not None:
                recs.append((title, "".join(lines).replace(" ", "").replace("\r", "")))
            title, lines = line[1:].rstrip(), []
        elif title is not None:
            lines.append(line.rstrip())
    if title is not None:
        recs.append((title, "".join(lines).replace(" ", "").replace("\r", "")))
    return recs


def ref_sort(inp, out):
    """readAndSortFasta, :150-166"""
    occ = {}
    with open(out, "w") as f:
        for d, s in sorted(simple_fasta(inp), key=lambda x: x[0]):
            f.write(">" + d + "\n" + s + "\n")
            occ[d] = occ.get(d, 0) + 1
    return occ


def ref_duplicate(reference, uncorrected, occ, new_unc, new_ref):
    """duplicateRefReads, :171-191"""
    with open(new_unc, "w") as nu, open(new_ref, "w") as nr:
        for unco, ref in zip(open(uncorrected).readlines(), open(reference).readlines()):
            if ">" not in ref:
                if header in occ:
                    for t in range(occ[header]):
                        nr.write(">" + header + "_" + str(t) + "\n" + ref.rstrip() + "\n")
                        nu.write(">" + header + "_" + str(t) + "\n" + unco.rstrip() + "\n")
            else:
                header = ref.rstrip()[1:]


def test_example_files_are_prepared_like_the_golden(tmp_path):
    from elector_b200 import prep
    from oracle import example_prep as ep
    src, w = str(tmp_path / "src"), str(tmp_path / "w")
    os.makedirs(src); os.makedirs(w)
    ep.unpack(src)
    # formatHeader(lordec, split) is `sed 's/_[0-9]*$//g'` on the corrected reads (readAndSortFiles.py:212): not part of this step
    with open(w + "/corrected_formatted.fa", "w") as f:
        for line in open(os.path.join(src, ep.FILES[2])):
            f.write(re.sub(r"_[0-9]*$", "", line.rstrip("\n")) + "\n")
    n_ref = prep.sort_fasta(os.path.join(src, ep.FILES[0]), w + "/reference_sorted.fa")
    n_unc = prep.sort_fasta(os.path.join(src, ep.FILES[1]), w + "/uncorrected_sorted.fa")
    n_cor = prep.sort_fasta(w + "/corrected_formatted.fa", w + "/corrected_sorted.fa")
    assert n_ref == n_unc
    n = prep.duplicate_reads(w + "/reference_sorted.fa", w + "/uncorrected_sorted.fa", w + "/corrected_sorted.fa", w + "/ref_dup.fa", w + "/unc_dup.fa")
    assert n == n_cor == 484
    g = load_example_golden()["sorted"]
    assert md5_file(w + "/ref_dup.fa") == g["ref"]
    assert md5_file(w + "/unc_dup.fa") == g["unc"]
    assert md5_file(w + "/corrected_sorted.fa") == g["cor"]


def test_fasta_shapes_against_the_restated_reference(tmp_path):
    from elector_b200 import prep
    d = str(tmp_path)
    open(d + "/ref.fa", "w", newline="").write("; a comment before the first record\n>r2 b\nACGT\nAC GT\n>r10\nAAAA\r\n>r1\n\nCC\n>zz only here\nGG\n>r2 b\nTTTT  \n")
    open(d + "/unc.fa", "w", newline="").write(">r2 b\nACGA\n>r10\nAAAT\n>r1\nCA\n>zz only here\nGT\n>r2 b\nTTTA\n")
    open(d + "/cor.fa", "w", newline="").write(">r2 b\nAC\n>r1\nCC\n>r2 b\nGT\n>r2 b\nTT\n>r10\nA\nA\n")
    for k in ("ref", "unc", "cor"):
        prep.sort_fasta("%s/%s.fa" % (d, k), "%s/%s_sorted.fa" % (d, k))
        occ = ref_sort("%s/%s.fa" % (d, k), "%s/%s_expect.fa" % (d, k))
        assert open("%s/%s_sorted.fa" % (d, k)).read() == open("%s/%s_expect.fa" % (d, k)).read(), k
    n = prep.duplicate_reads(d + "/ref_sorted.fa", d + "/unc_sorted.fa", d + "/cor_sorted.fa", d + "/ref_dup.fa", d + "/unc_dup.fa")
    ref_duplicate(d + "/ref_expect.fa", d + "/unc_expect.fa", occ, d + "/unc_dup_expect.fa", d + "/ref_dup_expect.fa")
    assert open(d + "/ref_dup.fa").read() == open(d + "/ref_dup_expect.fa").read()
    assert open(d + "/unc_dup.fa").read() == open(d + "/unc_dup_expect.fa").read()
    assert n == open(d + "/ref_dup.fa").read().count(">") == 8   # r1 x1, r10 x1, both r2 b records x3

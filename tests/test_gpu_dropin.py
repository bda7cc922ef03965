"""GPU: the drop-ins INSIDE ELECTOR.  oracle/_ref/elector_tree is an ELECTOR installation with the reference's own
unmodified Python (oracle/build_ref.sh lays it out; it travels to the GPU box like the reference binaries):
  - level 1, binary swap: the reference's alignment.getPOA run on a copy of that tree whose bin/poa and bin/masterSplitter are the CUDA
    executables -> msa.fa md5 == the golden (c91df333...) and the two counters;
  - level 3, elector_b200.alignment.getPOA / elector_b200.computeStats.outputRecallPrecision called like the reference's functions ->
    the same msa.fa bytes, return values, per_read_metrics.txt, read_size_distribution.txt, log and printed summary."""
import io
import os
import shutil
import subprocess
import sys
import contextlib

import pytest

from conftest import ROOT, md5_file

pytestmark = pytest.mark.gpu
TREE = os.path.join(ROOT, "oracle", "_ref", "elector_tree")


def swapped_tree(dst):
    """the reference installation with the two executables swapped (INTEGRATION.md level 1)"""
    from elector_b200.lib import library_path, poa_binary_path, splitter_binary_path
    shutil.copytree(TREE, dst)
    shutil.copy(poa_binary_path(), dst + "/bin/poa")
    shutil.copy(splitter_binary_path(), dst + "/bin/masterSplitter")
    shutil.copy(os.path.join(os.path.dirname(poa_binary_path()), "elector_server"), dst + "/bin/elector_server")   # the persistent service (csrc/service.h)
    shutil.copy(library_path(), dst + "/libelector_poa.so")      # the executables' rpath is $ORIGIN/..
    return dst


def run_reference_getpoa(tree, work, out, threads=4):
    code = ("import elector.alignment as a; r = a.getPOA(%r, %r, %r, %d, %r, 0.1); print('RESULT', r[0], r[1])"
            % (work + "/cor.fa", work + "/ref.fa", work + "/unc.fa", threads, out))
    p = subprocess.run([sys.executable, "-c", code], cwd=tree, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    t = p.stdout[p.stdout.index("RESULT"):].split()      # behind the 200 progress dashes of alignment.py:125
    return int(t[1]), int(t[2])


@pytest.mark.skipif(not os.path.isdir(TREE), reason="oracle/_ref/elector_tree not built")
def test_reference_getpoa_with_swapped_binaries(example_chain, tmp_path):
    g = example_chain["gold"]
    tree = swapped_tree(str(tmp_path / "tree"))
    out = str(tmp_path / "out")
    os.makedirs(out)
    assert run_reference_getpoa(tree, example_chain["work"], out) == (g["small_reads"], g["wrongly_cor_reads"])
    assert md5_file(out + "/msa.fa") == g["msa_md5"]


def test_inprocess_getpoa_and_report(example_chain, tmp_path):
    from elector_b200 import alignment, computeStats
    from test_report import check_against_golden
    g, work = example_chain["gold"], example_chain["work"]
    out = str(tmp_path / "out")
    os.makedirs(out)
    assert alignment.getPOA(work + "/cor.fa", work + "/ref.fa", work + "/unc.fa", 4, out, 0.1) == (g["small_reads"], g["wrongly_cor_reads"])
    assert md5_file(out + "/msa.fa") == g["msa_md5"]
    for reuse in (True, False):          # from the counters getPOA kept; from msa.fa alone (tally on the device again)
        if not reuse:
            alignment.LAST = None
        log, buf = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = computeStats.outputRecallPrecision(work + "/cor.fa", out, log, g["small_reads"], g["wrongly_cor_reads"], 5, 0.1, "read_size_distribution.txt", {},
                                                     0, 0, None, compensated_sum=True)
        check_against_golden({"log": log.getvalue(), "stdout": buf.getvalue(), "assessed_reads": ret[0], "trimmed_or_split": ret[17], "extended_reads": int(ret[12]),
                              "size_distribution_complete": 1}, out, g)
        assert ret[:6] == (459, 4454164, 0.9938972, 0.995006, 0.9938413, 1 - 0.9938413)
        assert ret[15] == [159839, 160207, 154842] and ret[16] == [8119, 8161, 12148] and ret[18] == 0.9925

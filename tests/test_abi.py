"""CPU: the C-ABI library loads, exports every symbol include/elector_poa.h declares, and
fails loudly (never falls back) when it cannot run on a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np

import pytest

from conftest import ROOT, has_gpu


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "elector_poa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elector_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import elector_b200
    lib = elector_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_matrix_errors_have_reference_exit_semantics(tmp_path):
    """unreadable matrix -> ELECTOR_EMATRIX (reference: exit 1, main.c:149-155); matrix outside the
    supported class -> ELECTOR_EUNSUPPORTED; both are decided before any device is touched."""
    import elector_b200
    from elector_b200.matrix import default_matrix_text
    with pytest.raises(elector_b200.ElectorError) as e:
        elector_b200.PoaContext(0, str(tmp_path / "nope.mat"))
    assert e.value.code == -2
    bad = tmp_path / "decay.mat"
    bad.write_text(default_matrix_text(gaps=(10, 5, 1)))
    with pytest.raises(elector_b200.ElectorError) as e:
        elector_b200.PoaContext(0, str(bad))
    assert e.value.code == -3
    bad2 = tmp_path / "trunc.mat"
    bad2.write_text("GAP-PENALTIES=10 5 5\n  a c\na 0 -1\nc -1\n")
    with pytest.raises(elector_b200.ElectorError) as e:
        elector_b200.PoaContext(0, str(bad2))
    assert e.value.code == -2


@pytest.mark.skipif(has_gpu(), reason="checks the no-device failure mode")
def test_no_device_is_an_error_not_a_fallback():
    import elector_b200
    with pytest.raises(elector_b200.ElectorError) as e:
        elector_b200.PoaContext(0)
    assert e.value.code == -4 and "no CPU path" in str(e.value)


def test_cli_usage_and_exit_codes(golden_dir, tmp_path):
    from elector_b200.lib import poa_binary_path
    exe = poa_binary_path()
    assert os.path.exists(exe), "build() must produce elector_b200/bin/poa"
    p = subprocess.run([exe], capture_output=True)
    assert p.returncode == 255 and b"Usage" in p.stderr           # exit(-1), main.c:40-83
    p = subprocess.run([exe, "-pathMatrix", str(tmp_path / "missing.mat"), "-pir", str(tmp_path / "o")], capture_output=True)
    assert p.returncode == 1                                       # main.c:149-155
    if not has_gpu():
        d = golden_dir
        p = subprocess.run([exe, "-pir", str(tmp_path / "o.pir"), "-corrected_reads_fasta", d + "/edge.cor.fa",
                            "-reference_reads_fasta", d + "/edge.ref.fa", "-uncorrected_reads_fasta", d + "/edge.unc.fa",
                            "-pathMatrix", d + "/blosum80.mat"], capture_output=True)
        assert p.returncode == 1 and b"no CPU path" in p.stderr


def test_product_never_references_the_oracle():
    """the oracle is test infrastructure: nothing under elector_b200/ or include/ may mention it"""
    for base in ("elector_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".c", ".inl")):
                    text = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle" not in text.lower() or f == "__init__.py" and False, "%s mentions the oracle" % f


def test_generated_matrix_is_the_shipped_one(tmp_path):
    from elector_b200 import write_default_matrix
    from oracle import oracle
    m = oracle.matrix(write_default_matrix(str(tmp_path / "m.mat")))
    d = oracle.matrix(None)
    assert m.nsymbol == d.nsymbol == 31 and m.symbol == d.symbol
    assert [list(r)[:31] for r in list(m.score)[:31]] == [list(r)[:31] for r in list(d.score)[:31]]
    assert list(m.gap_penalty_x)[:17] == [10] + [5] * 15 + [0]


def test_pack_letters_roundtrip():
    import elector_b200
    s = np.frombuffer(b"ACGTacgtNNRYKACGTTTGA" * 37 + b"n", np.uint8)
    p = elector_b200.pack_letters(s)
    assert p.n_letters == len(s)
    got = np.frombuffer(b"ACGT", np.uint8)[(p.bits[np.arange(len(s)) >> 2] >> (2 * (np.arange(len(s)) & 3))) & 3].copy()
    got[p.exc_pos] = p.exc_byte
    up = np.frombuffer(bytes(s).upper(), np.uint8)
    keep = np.ones(len(s), bool); keep[p.exc_pos] = False
    assert np.array_equal(got[keep], up[keep]) and np.array_equal(got[p.exc_pos], s[p.exc_pos])
    assert all(bytes([b]) not in b"ACGTacgt" for b in p.exc_byte)


def test_unpack_columns_is_the_inverse_of_the_column_code():
    """elector_unpack_columns (host helper of the one-byte-per-column wire format, include/elector_poa.h): byte = ref + 6 cor +
    36 unc over ELECTOR_COLUMN_CHARS; 255 (a column with a character outside the code) comes out as '?' in the three rows"""
    import elector_b200
    lib = elector_b200.load_library()
    chars = np.frombuffer(b".acgtn", np.uint8)
    rng = np.random.default_rng(3)
    n = 10007
    k = rng.integers(0, 6, size=(3, n))
    codes = (k[0] + 6 * k[1] + 36 * k[2]).astype(np.uint8)
    esc = rng.integers(0, n, 50)
    codes[esc] = 255
    out = [np.zeros(n, np.uint8) for _ in range(3)]
    lib.elector_unpack_columns(ctypes.c_void_p(codes.ctypes.data), ctypes.c_int64(n), *(ctypes.c_void_p(o.ctypes.data) for o in out))
    keep = np.ones(n, bool)
    keep[esc] = False
    for s in range(3):
        assert np.array_equal(out[s][keep], chars[k[s]][keep])
        assert (out[s][esc] == ord("?")).all()

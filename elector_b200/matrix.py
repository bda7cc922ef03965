"""The score matrix ELECTOR always passes to poa (elector/alignment.py:60 -> blosum80.mat):
identity 0 / -10 over the 31-symbol alphabet, gap penalties 10 5 5, truncation 10, decay 5.
Generated rather than shipped, in the reference's file format (seq_util.c:82-217)."""

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX?agtcu]n"


def default_matrix_text(match=0, mismatch=-10, gaps=(10, 5, 5), trunc=10, decay=5, alphabet=ALPHABET):
    lines = ["# identity matrix in the format of poaV2 score files",
             "GAP-TRUNCATION-LENGTH=%d" % trunc,
             "GAP-DECAY-LENGTH=%d" % decay,
             "GAP-PENALTIES=%d %d %d" % tuple(gaps),
             "  " + " ".join(alphabet)]
    for i, a in enumerate(alphabet):
        lines.append(a + " " + " ".join(str(match if i == j else mismatch) for j in range(len(alphabet))))
    return "\n".join(lines) + "\n"


def write_default_matrix(path, **kw):
    with open(path, "w") as f:
        f.write(default_matrix_text(**kw))
    return path

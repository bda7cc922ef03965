// host_io.hpp -- host-side parsing for the `poa` drop-in: score-matrix file and FASTA
// shards.  Observable behaviour follows the reference readers; structure is ours.
//   score matrix : src/poa-graph/seq_util.c:82-217 (read_score_matrix)
//   FASTA        : src/poa-graph/fasta_format.c:10-66 (read_fasta), create_seq.c:22-51
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace elector {

struct ScoreMatrix {
  int nsymbol = 0;
  std::string symbol;               // alphabet, in file order
  std::vector<int> score;           // nsymbol x nsymbol, score[row*nsymbol+col]
  int gap_open[2] = {12, 12};       // [0] = x set, [1] = y set (GAP-PENALTIES-X overrides [1])
  int gap_a1[2] = {2, 2};
  int gap_a2[2] = {0, 0};
  int trunc_len = 10, decay_len = 5;
  std::vector<int> pen_x, pen_y;    // 0..max_gap+1

  void build_tables() {
    const int T = trunc_len, D = decay_len, M = T + D;
    pen_x.assign(M + 2, 0);
    pen_y.assign(M + 2, 0);
    pen_x[0] = gap_open[0];
    pen_y[0] = gap_open[1];
    for (int i = 1; i < T; ++i) { pen_x[i] = gap_a1[0]; pen_y[i] = gap_a1[1]; }
    for (int i = 0; i < D; ++i) {
      const double dx = (gap_a1[0] - gap_a2[0]) / double(D + 1), dy = (gap_a1[1] - gap_a2[1]) / double(D + 1);
      pen_x[T + i] = int(gap_a1[0] - (i + 1) * dx);
      pen_y[T + i] = int(gap_a1[1] - (i + 1) * dy);
    }
    pen_x[M] = gap_a2[0];
    pen_y[M] = gap_a2[1];
  }

  void set_default() {  // the values of the matrix file ELECTOR always passes (alignment.py:60)
    symbol = "ARNDCQEGHILKMFPSTWYVBZX?agtcu]n";
    nsymbol = (int)symbol.size();
    score.assign((size_t)nsymbol * nsymbol, -10);
    for (int i = 0; i < nsymbol; ++i) score[(size_t)i * nsymbol + i] = 0;
    gap_open[0] = gap_open[1] = 10;
    gap_a1[0] = gap_a1[1] = 5;
    gap_a2[0] = gap_a2[1] = 5;
    trunc_len = 10;
    decay_len = 5;
    build_tables();
  }

  // returns nsymbol (>0) or <=0 on failure, like read_score_matrix
  int load(const char *path) {
    FILE *f = std::fopen(path, "r");
    if (!f) return -2;
    char line[1024];
    bool expect_symbols = true;
    std::vector<std::vector<int>> rows;
    std::vector<int> row_of;
    symbol.clear();
    while (std::fgets(line, 1023, f)) {
      int a, b, c;
      if (line[0] == '#' || line[0] == '\n') continue;
      if (std::sscanf(line, "GAP-TRUNCATION-LENGTH=%d", &a) == 1) { trunc_len = a; continue; }
      if (std::sscanf(line, "GAP-DECAY-LENGTH=%d", &a) == 1) { decay_len = a; continue; }
      if (std::sscanf(line, "GAP-PENALTIES=%d %d %d", &a, &b, &c) == 3) {
        gap_open[0] = gap_open[1] = a; gap_a1[0] = gap_a1[1] = b; gap_a2[0] = gap_a2[1] = c;
        continue;
      }
      if (std::sscanf(line, "GAP-PENALTIES-X=%d %d %d", &a, &b, &c) == 3) {
        gap_open[1] = a; gap_a1[1] = b; gap_a2[1] = c;
        continue;
      }
      if (expect_symbols) {  // the alphabet line (a later unmatched line restarts it, as in the reference)
        for (const char *p = line; *p; ++p)
          if (!std::isspace((unsigned char)*p)) symbol.push_back(*p);
        expect_symbols = false;
        continue;
      }
      const size_t pos = symbol.rfind(line[0]);  // the reference scans from the end
      if (pos == std::string::npos) { std::fclose(f); return -1; }
      std::vector<int> vals(symbol.size());
      int off = 1, used = 0;
      for (size_t i = 0; i < symbol.size(); ++i) {
        if (std::sscanf(line + off, "%d%n", &vals[i], &used) != 1) { std::fclose(f); return -1; }
        off += used;
      }
      rows.push_back(vals);
      row_of.push_back((int)pos);
    }
    std::fclose(f);
    nsymbol = (int)symbol.size();
    if (nsymbol <= 0) return 0;
    score.assign((size_t)nsymbol * nsymbol, 0);
    for (size_t k = 0; k < rows.size(); ++k)
      for (int i = 0; i < nsymbol && i < (int)rows[k].size(); ++i) score[(size_t)row_of[k] * nsymbol + i] = rows[k][i];
    build_tables();
    return nsymbol;
  }

  // byte -> symbol index after lower-casing (create_seq.c:39-43), limit_residues
  // (seq_util.c:253-263: unknown -> symbol[0]) and index_symbols (:37-52: last match wins)
  int code_of(int byte) const {
    int c = std::tolower(byte & 0xff);
    if (c == 0 || symbol.find((char)c) == std::string::npos) c = (unsigned char)symbol[0];
    const size_t pos = symbol.rfind((char)c);
    return pos == std::string::npos ? nsymbol - 1 : (int)pos;
  }
};

struct FastaRecord {
  std::string name, title;
  int64_t off = 0;
  int32_t len = 0;
};

struct FastaFile {
  std::vector<FastaRecord> rec;
  std::string seq;  // all sequences concatenated (raw letters, whitespace stripped)
  std::vector<int64_t> offsets() const {
    std::vector<int64_t> o(rec.size() + 1, 0);
    for (size_t i = 0; i < rec.size(); ++i) { o[i] = rec[i].off; o[i + 1] = rec[i].off + rec[i].len; }
    return o;
  }
};

// Reads a FASTA shard with the reference's quirks: 32 KiB line chunks, '#' comment and
// '*' lines skipped, records with an empty sequence dropped, reading stops at a '#' line
// once a record exists, name = first token (cut to 511 chars), title = rest or "untitled".
inline int read_fasta_file(const char *path, FastaFile &out) {
  FILE *f = std::fopen(path, "r");
  if (!f) return -1;
  static const int CHUNK = 32768;
  std::vector<char> line(CHUNK), name(CHUNK + 8), title(CHUNK + 8);
  std::string cur;
  bool have_name = false;
  auto commit = [&]() {
    if (!have_name) return;
    std::string s;
    s.reserve(cur.size());
    for (char ch : cur)
      if (!std::isspace((unsigned char)ch)) s.push_back(ch);
    // the reference tests the raw buffer for emptiness before stripping (fasta_format.c:30)
    if (cur.empty()) return;
    FastaRecord r;
    r.name.assign(name.data());
    if (r.name.size() > 511) r.name.resize(511);
    r.title.assign(title.data());
    r.off = (int64_t)out.seq.size();
    r.len = (int32_t)s.size();
    out.seq += s;
    out.rec.push_back(r);
  };
  while (std::fgets(line.data(), CHUNK - 1, f)) {
    if (char *nl = std::strrchr(line.data(), '\n')) *nl = '\0';
    switch (line[0]) {
      case '#': break;
      case '>':
        commit();
        name[0] = '\0';
        if (std::sscanf(line.data() + 1, "%s %[^\n]", name.data(), title.data()) < 2) std::strcpy(title.data(), "untitled");
        have_name = name[0] != '\0';
        cur.clear();
        break;
      case '*': break;
      default:
        if (have_name) cur += line.data();
    }
    const int c = std::getc(f);
    if (c == EOF) break;
    std::ungetc(c, f);
    if (c == '#' && !out.rec.empty()) break;
  }
  commit();
  std::fclose(f);
  return (int)out.rec.size();
}

}  // namespace elector

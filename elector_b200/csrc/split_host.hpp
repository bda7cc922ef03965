// split_host.hpp -- host side of the window cutting: the reference's `masterSplitter` command line, its reading of the three
// read files and its output files (Master_Splitter.cpp main(), :352-472), shared by the drop-in executable
// (csrc/splitter_main.cpp over the C-ABI) and the CPU emulation harness (tests/emul/split_emul.cu).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "split_kernel.cuh"

namespace elector {

// largest_fragment() of the reference (:158-169) on the text "header\nseq\n..." that a record list stands for: the first
// line counts its own length, every later line one more (the scan measures from newline to newline)
inline unsigned split_largest_fragment(const SplitWin *w, int n, int header_len) {
  unsigned res = (unsigned)header_len + (n > 1 ? 1u : 0u);
  for (int i = 0; i < n; ++i) res = std::max(res, (unsigned)w[i].rn + 1u);
  return res;
}

struct SplitChoice { int status = 0, k = 15; };   // status 0: cut, 1: corrected read too short (small_reads), 2: not cut (wrongly_cor_reads)

// the triplets of one round: letters of the three kinds back to back, with offsets
struct SplitBatch {
  std::vector<std::string> header;
  std::vector<uint8_t> letters[3];
  std::vector<int64_t> off[3] = {{0}, {0}, {0}};
  size_t n() const { return header.size(); }
  const uint8_t *seq(int kind, size_t t) const { return letters[kind].data() + off[kind][t]; }
  int len(int kind, size_t t) const { return (int)(off[kind][t + 1] - off[kind][t]); }
  int header_len(size_t t) const { return (int)header[t].size(); }
  void add(const std::string &h, const std::string &r, const std::string &a, const std::string &b) {
    header.push_back(h);
    const std::string *s[3] = {&r, &a, &b};
    for (int k = 0; k < 3; ++k) { letters[k].insert(letters[k].end(), s[k]->begin(), s[k]->end()); off[k].push_back((int64_t)letters[k].size()); }
  }
};

struct HostScratch {   // scratch of one job for the host build of the kernels' code, laid out like the kernel lays it out
  std::vector<uint32_t> mem;
  SplitScratch sc, sc2;
  bool tight;
  // longest: the longest read of any kind; tight: the smallest table the kernel accepts (the fullest it ever runs)
  explicit HostScratch(size_t longest, bool tight_ = false) : tight(tight_) {
    const int n = (int)longest;
    mem.resize((size_t)(split_fixed_words(n, n, n, split_anchor_bound(n, 20), sub_anchors(n)) + 2ull * longest + 64) + 4);
    fit(n, n, n);
  }
  static int32_t sub_anchors(int longest) { return (int32_t)((longest / 8 + 16 + 3) & ~3); }
  // the arrays of one job, like split_jobs_kernel places them in its pool
  void fit(int nr, int na, int nb) {
    const int32_t ma = split_anchor_bound(nr, 20), sa = sub_anchors(nr > na ? nr : na);
    uint32_t *base = mem.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    uint64_t words = mem.size() - 4;
    if (tight) words = split_fixed_words(nr, na, nb, ma, sa) + split_min_slots(nr);
    if (!split_carve(sc, sc2, base, words, nr, na, nb, ma, sa)) abort();
  }
};

// the reference's command line and files
struct SplitCli {
  std::string in[3], out[3], out_dir;
  int k = 7, nb_file = 200;
  uint64_t max_amount = 10000;
  double threshold = 0.1;
  uint64_t pos[3] = {0, 0, 0};
  bool eof[3] = {false, false, false};

  bool parse(int argc, char **argv) {
    if (argc < 12) return false;
    for (int i = 0; i < 3; ++i) { in[i] = argv[1 + i]; out[i] = argv[4 + i]; }
    k = atoi(argv[7]); nb_file = atoi(argv[8]); max_amount = (uint64_t)atoi(argv[9]); threshold = atof(argv[10]); out_dir = argv[11];
    return nb_file > 0;
  }
  // Reads the triplets of one round like main() (:396-446): resumes at progress.txt, two lines per record, stops after the
  // triplet whose index exceeds max_amount or at the end of a file; records whose reference has at most 2 letters are
  // dropped without being counted.  Returns 0 = all files read to the end ... 1 = more to come, -1 error.
  int read_round(SplitBatch &b) {
    std::ifstream f[3], prog(out_dir + "/progress.txt");
    for (int i = 0; i < 3; ++i) f[i].open(in[i]);
    if (prog.good() && !prog.eof()) {
      std::string line;
      for (int i = 0; i < 3; ++i) { std::getline(prog, line); pos[i] = (uint64_t)atoll(line.c_str()); f[i].seekg((std::streamoff)pos[i], f[i].beg); }
    }
    uint64_t i = 0;
    std::string h[3], s[3];
    bool stop = false;
    while (!f[0].eof() && !f[1].eof() && !f[2].eof() && !stop) {
      if (i > max_amount) break;
      for (unsigned ii = 0; ii < 1000; ++ii) {
        if (i > max_amount) continue;
        for (int q = 0; q < 3; ++q) { std::getline(f[q], h[q]); std::getline(f[q], s[q]); }
        if (s[0].size() > 2) {
          b.add(h[0], s[0], s[1], s[2]);
          ++i;
          for (int q = 0; q < 3; ++q) { h[q].clear(); s[q].clear(); }   // (:439) a getline on a finished stream leaves its string as it was
        }
      }
    }
    for (int q = 0; q < 3; ++q) { eof[q] = f[q].eof(); if (!eof[q]) pos[q] = (uint64_t)f[q].tellg(); }
    return (eof[0] || eof[1] || eof[2]) ? 0 : 1;
  }
  // Writes the round's shard files, the two counters and progress.txt (:389-393,:447-471); returns the reference's exit code.
  // status[t]: SplitChoice::status; count(t) = records of a cut triplet; get(t, i, q, &ptr, &len) = letters of record i, kind q.
  template <class Count, class Get>
  int write_round(const SplitBatch &b, const int32_t *status, Count count, Get get, int more) const {
    const int64_t factor = (int64_t)(max_amount / (uint64_t)nb_file) + 1;
    printf("%llu %d %lld\n", (unsigned long long)max_amount, nb_file, (long long)(factor - 1));   // (:367, before factor += 1)
    std::vector<std::string> text[3];
    for (int q = 0; q < 3; ++q) text[q].resize((size_t)nb_file);
    int small_reads = 0, wrong_reads = 0;
    for (size_t t = 0; t < b.n(); ++t) {
      const size_t shard = (size_t)((int64_t)t / factor);
      if (shard >= (size_t)nb_file) break;
      const std::string &h = b.header[t];
      if (status[t] != 0) {
        for (int q = 0; q < 3; ++q) { text[q][shard] += h; text[q][shard] += "\nAAA\n"; }
        if (status[t] == 1) ++small_reads; else ++wrong_reads;
        continue;
      }
      const int64_t nrec = count(t);
      for (int64_t i = 0; i < nrec; ++i)
        for (int q = 0; q < 3; ++q) {
          const char *p = nullptr; size_t len = 0;
          get(t, i, q, &p, &len);
          text[q][shard] += h; text[q][shard] += '\n';
          text[q][shard].append(p, len);
          text[q][shard] += '\n';
        }
    }
    for (int q = 0; q < 3; ++q)
      for (int i = 0; i < nb_file; ++i) {
        std::ofstream o(out[q] + std::to_string(i), std::ofstream::trunc);
        o << text[q][(size_t)i];
      }
    { std::ofstream o(out_dir + "/small_reads.txt"); o << small_reads << std::endl; }
    { std::ofstream o(out_dir + "/wrongly_cor_reads.txt"); o << wrong_reads << std::endl; }
    // (:460-463 means to remove progress.txt at the end of the input, but hands remove() the buffer of a destroyed temporary:
    // with glibc the file stays, and so it does here)
    if (!more) return 0;
    std::ofstream o(out_dir + "/progress.txt");
    o << pos[0] << "\n" << pos[1] << "\n" << pos[2] << "\n" << std::flush;
    return 1;
  }
};

}  // namespace elector

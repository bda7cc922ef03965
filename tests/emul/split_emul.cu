// TEST INFRASTRUCTURE ONLY -- runs the window cutting of elector_b200/csrc/split_kernel.cuh serially on the CPU (the code is
// host/device dual) with the reference's own driver logic around it (Master_Splitter.cpp main(), :352-472: one round of at
// most 10 001 triplets, shard files by triplet index, the two counters), so that its output files can be compared byte for
// byte with those of the compiled reference (oracle/_ref/masterSplitter) in the GPU-less container.
// usage: split_emul REF.fa UNC.fa COR.fa OUT1 OUT2 OUT3 k nb_file max_amount threshold OUTDIR      (the reference's arguments)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../../elector_b200/csrc/split_kernel.cuh"
#include "../../elector_b200/csrc/split_host.hpp"

using namespace elector;

int main(int argc, char **argv) {
  if (argc < 12) { fprintf(stderr, "usage: %s REF UNC COR OUT1 OUT2 OUT3 k nb_file max_amount threshold OUTDIR\n", argv[0]); return 2; }
  SplitCli cli;
  if (!cli.parse(argc, argv)) return 2;
  SplitBatch batch;
  const int rc_read = cli.read_round(batch);
  if (rc_read < 0) return 2;
  // every job on the CPU, one after the other
  std::vector<SplitChoice> choice(batch.n());
  std::vector<std::vector<SplitWin>> wins(batch.n());
  size_t longest = 0;
  for (size_t t = 0; t < batch.n(); ++t) for (int q = 0; q < 3; ++q) longest = std::max(longest, (size_t)batch.len(q, t));
  HostScratch hs(longest, getenv("SPLIT_EMUL_TIGHT") != nullptr);
  // SPLIT_EMUL_PACKED: the reads come from letters packed once at 2 bits each, like split_prepack_kernel leaves them on the device
  std::vector<uint32_t> packed[3];
  if (getenv("SPLIT_EMUL_PACKED"))
    for (int q = 0; q < 3; ++q) {
      packed[q].assign(batch.letters[q].size() / 16 + 3, 0u);
      for (size_t t = 0; t < batch.letters[q].size(); ++t) {
        const uint8_t c = batch.letters[q][t];
        packed[q][t >> 4] |= (c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u) << (2 * (t & 15));
      }
    }
  for (size_t t = 0; t < batch.n(); ++t) {
    SplitSeq ref{batch.seq(0, t), batch.len(0, t)}, S1{batch.seq(1, t), batch.len(1, t)}, S2{batch.seq(2, t), batch.len(2, t)};
    if (!packed[0].empty()) {
      ref.pk = packed[0].data(); ref.g = batch.off[0][t]; S1.pk = packed[1].data(); S1.g = batch.off[1][t]; S2.pk = packed[2].data(); S2.g = batch.off[2][t];
    }
    choice[t].status = 0;
    if (!((double)S2.n / ref.n >= cli.threshold)) { choice[t].status = 1; continue; }
    hs.fit(ref.n, S1.n, std::max(S2.n, ref.n));
    std::vector<SplitWin> best;
    unsigned best_largest = 0;
    for (int ki = 0; ki < 4; ++ki) {
      const int k = 15 - 2 * ki;
      std::vector<SplitWin> out((size_t)ref.n / 8 + 16);
      int s_int[16];
      const int n = split_job(hs.sc, hs.sc2, ref, S1, S2, k, out.data(), (int)out.size(), s_int);
      if (n < 0) { fprintf(stderr, "record capacity\n"); return 3; }
      out.resize(n);
      const unsigned largest = split_largest_fragment(out.data(), n, batch.header_len(t));
      if (ki == 0 || largest < best_largest) { best.swap(out); best_largest = largest; choice[t].k = k; }
      else break;
    }
    if (best.size() <= 1) choice[t].status = 2;
    else wins[t] = best;
  }
  std::vector<int32_t> status(batch.n());
  for (size_t t = 0; t < batch.n(); ++t) status[t] = choice[t].status;
  return cli.write_round(batch, status.data(), [&](size_t t) { return (int64_t)wins[t].size(); },
                         [&](size_t t, int64_t i, int q, const char **p, size_t *len) {
                           const SplitWin &w = wins[t][(size_t)i];
                           const int st[3] = {w.r0, w.a0, w.b0}, ln[3] = {w.rn, w.an, w.bn};
                           if (st[q] < 0) { *p = "N"; *len = 1; }
                           else { *p = reinterpret_cast<const char *>(batch.seq(q, t)) + st[q]; *len = (size_t)ln[q]; }
                         }, rc_read);
}

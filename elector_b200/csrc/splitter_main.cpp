// masterSplitter -- drop-in for the reference's window cutter (src/split/Master_Splitter.cpp main(), :352-472) over the C-ABI:
// same command line (elector/alignment.py:99), same shard files out1<i> / out2<i> / out3<i>, small_reads.txt,
// wrongly_cor_reads.txt, progress.txt and exit code (1 = more rounds to come).  The cutting itself is elector_split_run
// (split_kernel.cuh on the device); this file is the reference's file handling around it (split_host.hpp).
//   masterSplitter REF.fa UNC.fa COR.fa OUT1 OUT2 OUT3 k nb_file max_amount threshold OUTDIR
// ELECTOR_DEVICE selects the CUDA ordinal (default 0).  There is no CPU path: without a device the program exits with 3.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/elector_poa.h"
#include "split_host.hpp"
#ifndef ELECTOR_SERVER
#include "service.h"
#endif

using namespace elector;

int main(int argc, char **argv) {
  SplitCli cli;
#ifndef ELECTOR_SERVER
  { int code; if (argc >= 12 && svc_try_call(SVC_KIND_SPLITTER, argc, argv, &code)) return code; }   // the persistent service (service.h)
#endif
  if (!cli.parse(argc, argv)) { fprintf(stderr, "usage: %s REF UNC COR OUT1 OUT2 OUT3 k nb_file max_amount threshold OUTDIR\n", argv[0]); return 2; }
  SplitBatch b;
  const int more = cli.read_round(b);
  const int64_t n = (int64_t)b.n();
  std::vector<int32_t> status((size_t)n), k_used((size_t)n), hl((size_t)n);
  std::vector<int64_t> rf((size_t)n + 1, 0), wo[3];
  std::vector<char> wl[3];
  if (n > 0) {
    elector_ctx *ctx = nullptr;
    const char *dev = getenv("ELECTOR_DEVICE");
    if (elector_poa_init(dev ? atoi(dev) : 0, nullptr, &ctx) != ELECTOR_OK) { fprintf(stderr, "masterSplitter: %s\n", elector_last_error(nullptr)); return 3; }
    for (int64_t t = 0; t < n; ++t) hl[(size_t)t] = b.header_len((size_t)t);
    int64_t cap_w = 0, cap[3] = {0, 0, 0};
    elector_split_bounds(n, b.off[0].data(), b.off[1].data(), b.off[2].data(), &cap_w, &cap[0], &cap[1], &cap[2]);
    for (int q = 0; q < 3; ++q) { wo[q].resize((size_t)cap_w + 1); wl[q].resize((size_t)cap[q]); }
    int64_t nw = 0;
    const int rc = elector_split_run(ctx, n, reinterpret_cast<const char *>(b.letters[0].data()), b.off[0].data(), reinterpret_cast<const char *>(b.letters[1].data()),
                                     b.off[1].data(), reinterpret_cast<const char *>(b.letters[2].data()), b.off[2].data(), hl.data(), cli.threshold, status.data(),
                                     k_used.data(), rf.data(), cap_w, wo[0].data(), wo[1].data(), wo[2].data(), wl[0].data(), cap[0], wl[1].data(), cap[1], wl[2].data(),
                                     cap[2], &nw);
    if (rc != ELECTOR_OK) { fprintf(stderr, "masterSplitter: %s\n", elector_last_error(ctx)); elector_poa_free(ctx); return 3; }
    elector_poa_free(ctx);
  }
  return cli.write_round(b, status.data(), [&](size_t t) { return rf[t + 1] - rf[t]; },
                         [&](size_t t, int64_t i, int q, const char **p, size_t *len) {
                           const int64_t w = rf[t] + i;
                           *p = wl[q].data() + wo[q][(size_t)w];
                           *len = (size_t)(wo[q][(size_t)w + 1] - wo[q][(size_t)w]);
                         }, more);
}

// split_capi.inl -- C-ABI entry points of the window cutting (included by capi.cu)
namespace {
struct SplitBufs {
  DevBuf let[3], pk[3], off[3], order, hl, st_in, pool, wins, win_off, job_n, job_l, status, kidx, nrec, nl[3], base[4], w_off[3], w_let[3], rf, misc;
};
SplitBufs *split_bufs(elector_ctx *ctx) {
  if (!ctx->split_state) ctx->split_state = new SplitBufs();
  return static_cast<SplitBufs *>(ctx->split_state);
}
}  // namespace

extern "C" {

int elector_split_bounds(int64_t n, const int64_t *ref_off, const int64_t *unc_off, const int64_t *cor_off, int64_t *win_cap, int64_t *ref_cap,
                         int64_t *unc_cap, int64_t *cor_cap) {
  if (n < 0 || (n > 0 && (!ref_off || !unc_off || !cor_off))) return ELECTOR_EINVAL;
  int64_t w = 0;
  for (int64_t t = 0; t < n; ++t) w += (ref_off[t + 1] - ref_off[t]) / 16 + 16;   // anchors are more than 20 letters apart
  if (win_cap) *win_cap = w;
  if (ref_cap) *ref_cap = (n ? ref_off[n] - ref_off[0] : 0) + 3 * n + 16;
  if (unc_cap) *unc_cap = (n ? unc_off[n] - unc_off[0] : 0) + 3 * n + 16;
  if (cor_cap) *cor_cap = (n ? cor_off[n] - cor_off[0] : 0) + 3 * n + w + 16;        // + one N per record at most
  return ELECTOR_OK;
}

void elector_split_release(elector_ctx *ctx) {
  if (!ctx || !ctx->split_state) return;
  SplitBufs *b = static_cast<SplitBufs *>(ctx->split_state);
  for (DevBuf *d : {&b->let[0], &b->let[1], &b->let[2], &b->pk[0], &b->pk[1], &b->pk[2], &b->off[0], &b->off[1], &b->off[2], &b->order, &b->hl, &b->st_in, &b->pool, &b->wins, &b->win_off, &b->job_n, &b->job_l,
                    &b->status, &b->kidx, &b->nrec, &b->nl[0], &b->nl[1], &b->nl[2], &b->base[0], &b->base[1], &b->base[2], &b->base[3], &b->w_off[0], &b->w_off[1],
                    &b->w_off[2], &b->w_let[0], &b->w_let[1], &b->w_let[2], &b->rf, &b->misc})
    d->release();
  delete b;
  ctx->split_state = nullptr;
}

}  // extern "C"

namespace {
// One round of window cutting with the result left on the device: b.w_off[k] / b.w_let[k] (k = 0 ref, 1 unc, 2 cor), b.rf
// (first window of every triplet), b.status, b.kidx.  tot[0] = windows, tot[1..3] = letters of the three kinds.
int split_device(elector_ctx *ctx, int64_t n, const char *ref, const int64_t *ref_off, const char *unc, const int64_t *unc_off, const char *cor,
                 const int64_t *cor_off, const int32_t *header_len, double threshold, int64_t tot[4]) {
  if (n > 0x1fffffff) return ctx->fail(ELECTOR_EINVAL, "too many triplets in one call");
  SplitBufs &b = *split_bufs(ctx);
  cudaStream_t st = ctx->stream;
  const char *h_let[3] = {ref, unc, cor};
  const int64_t *h_off[3] = {ref_off, unc_off, cor_off};
  // the host decides what main() decides before best_split (:414-415,:425-431): a corrected read shorter than the threshold share
  // of its reference is not cut; the longest reference read sizes the tables
  std::vector<int32_t> st_in((size_t)n), order((size_t)n);
  std::vector<int64_t> win_off((size_t)n + 1);
  // shared-memory shapes of the cutting kernel: two CTAs of 512 threads per SM with 112 KB each, or one of 1024 threads with 227 KB
  const uint32_t smem_a = 112u * 1024u / 4u, smem_b = (227u * 1024u - 256u) / 4u;
  int64_t longest = 0, fit_a = 0;
  uint64_t need_pool = 0;
  win_off[0] = 0;
  auto need_words = [](int64_t lr, int64_t la, int64_t lb, int32_t sub, uint64_t slots) {
    const int32_t ma = split_anchor_bound((int)lr, 20);
    return split_fixed_words((int)lr, (int)la, (int)std::max(lb, lr), ma, sub ? sub : ma) + slots;
  };
  for (int64_t t = 0; t < n; ++t) {
    const int64_t lr = ref_off[t + 1] - ref_off[t], la = unc_off[t + 1] - unc_off[t], lb = cor_off[t + 1] - cor_off[t];
    if (lr <= 0) return ctx->fail(ELECTOR_EINVAL, "triplet %lld has an empty reference read", (long long)t);
    if (lr > kSplitMaxRead || la > kSplitMaxRead || lb > kSplitMaxRead) return ctx->fail(ELECTOR_ETOOLARGE, "triplet %lld: read too long", (long long)t);
    st_in[(size_t)t] = ((double)lb / (double)lr >= threshold) ? 0 : 1;
    longest = std::max<int64_t>(longest, std::max<int64_t>(lr, std::max<int64_t>(la, lb)));
    if (need_words(lr, la, lb, 0, split_min_slots((int)lr)) <= smem_a) ++fit_a;
    win_off[(size_t)t + 1] = win_off[(size_t)t] + 4 * (lr / 16 + 16);
  }
  const bool shape_a = fit_a * 10 >= n * 9;
  const uint32_t smem_words = shape_a ? smem_a : smem_b;
  const int32_t sub_anchors = (int32_t)((longest / 8 + 16 + 3) & ~(int64_t)3);
  for (int64_t t = 0; t < n; ++t) {   // the pool holds any job of the call with two table slots per k-mer: what does not fit the shared memory runs there
    const int64_t lr = ref_off[t + 1] - ref_off[t], la = unc_off[t + 1] - unc_off[t], lb = cor_off[t + 1] - cor_off[t];
    if (!st_in[(size_t)t]) need_pool = std::max<uint64_t>(need_pool, need_words(lr, la, lb, sub_anchors, (uint64_t)split_min_slots((int)lr) + (uint64_t)lr / 4));   // ~1.55 slots per k-mer: the tables of all CTAs stay in the L2
  }
  {   // the longest reads first: a counting sort by length / 256
    const size_t nb = (size_t)(longest / 256 + 2);
    std::vector<int64_t> start(nb + 1, 0);
    for (int64_t t = 0; t < n; ++t) ++start[nb - 1 - (size_t)((ref_off[t + 1] - ref_off[t]) / 256)];
    int64_t acc = 0;
    for (size_t i = 0; i <= nb; ++i) { const int64_t c = start[i]; start[i] = acc; acc += c; }
    for (int64_t t = 0; t < n; ++t) order[(size_t)start[nb - 1 - (size_t)((ref_off[t + 1] - ref_off[t]) / 256)]++] = (int32_t)t;
  }
  const uint64_t cta_words = (need_pool + 31) & ~31ull;
  const int threads = shape_a ? 512 : 1024;
  int grid = (int)std::min<int64_t>(4 * n, (int64_t)ctx->sm_count * (shape_a ? 2 : 1));
  for (int k = 0; k < 3; ++k) {
    CU(b.let[k].reserve((size_t)(h_off[k][n] - h_off[k][0]) + 16)); CU(b.off[k].reserve((size_t)(n + 1) * 8));
    CU(cudaMemcpyAsync(b.let[k].p, h_let[k] + h_off[k][0], (size_t)(h_off[k][n] - h_off[k][0]), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b.off[k].p, h_off[k], (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  }
  CU(b.order.reserve((size_t)n * 4));
  CU(cudaMemcpyAsync(b.order.p, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CU(b.hl.reserve((size_t)n * 4)); CU(b.st_in.reserve((size_t)n * 4)); CU(b.win_off.reserve((size_t)(n + 1) * 8));
  CU(b.pool.reserve((size_t)cta_words * 4 * (size_t)grid));
  CU(b.wins.reserve((size_t)win_off[(size_t)n] * sizeof(SplitWin)));
  CU(b.job_n.reserve((size_t)n * 16)); CU(b.job_l.reserve((size_t)n * 16));
  CU(b.status.reserve((size_t)n * 4)); CU(b.kidx.reserve((size_t)n * 4)); CU(b.nrec.reserve((size_t)(n + 1) * 8));
  for (int k = 0; k < 3; ++k) CU(b.nl[k].reserve((size_t)(n + 1) * 8));
  for (int k = 0; k < 4; ++k) CU(b.base[k].reserve((size_t)(n + 1) * 8));
  CU(b.misc.reserve(64));
  CU(cudaMemcpyAsync(b.hl.p, header_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(b.st_in.p, st_in.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(b.win_off.p, win_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(b.misc.p, 0, 64, st));
  SplitArgs a;
  a.n_triplets = n;
  for (int k = 0; k < 3; ++k) { a.let[k] = b.let[k].as<uint8_t>() - h_off[k][0]; a.off[k] = b.off[k].as<int64_t>(); }
  a.header_len = b.hl.as<int32_t>(); a.status_in = b.st_in.as<int32_t>();
  a.order = b.order.as<int32_t>();
  for (int k = 0; k < 3; ++k) {   // the letters once at 2 bits each: the four jobs of a triplet read these
    const int64_t nl = h_off[k][n] - h_off[k][0], nwords = nl / 16 + 3;
    CU(b.pk[k].reserve((size_t)nwords * 4));
    split_prepack_kernel<<<(unsigned)std::min<int64_t>((nwords + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(b.let[k].as<uint8_t>(), nl, b.pk[k].as<uint32_t>(), nwords);
    a.pk[k] = b.pk[k].as<uint32_t>(); a.base[k] = h_off[k][0];
  }
  a.pool = b.pool.as<uint32_t>(); a.cta_words = cta_words; a.sub_anchors = sub_anchors; a.smem_words = smem_words;
  a.wins = b.wins.as<SplitWin>(); a.win_off = b.win_off.as<int64_t>();
  a.job_n = b.job_n.as<int32_t>(); a.job_largest = b.job_l.as<uint32_t>();
  a.counter = b.misc.as<int32_t>();
  CU(cudaEventRecord(ctx->ev_split0, st));
  CU(cudaFuncSetAttribute(split_jobs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_b * 4)));
  split_jobs_kernel<<<grid, threads, (size_t)smem_words * 4, st>>>(a);
  split_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a.status_in, a.job_n, a.job_largest, a.wins, a.win_off, b.status.as<int32_t>(), b.kidx.as<int32_t>(),
                                                                  b.nrec.as<int64_t>(), b.nl[0].as<int64_t>(), b.nl[1].as<int64_t>(), b.nl[2].as<int64_t>(), b.misc.as<int32_t>() + 1);
  scan_offsets_kernel<<<1, 1024, 0, st>>>(n, b.nrec.as<int64_t>(), b.base[0].as<int64_t>());
  for (int k = 0; k < 3; ++k) scan_offsets_kernel<<<1, 1024, 0, st>>>(n, b.nl[k].as<int64_t>(), b.base[k + 1].as<int64_t>());
  CU(cudaGetLastError());
#ifdef SPLIT_TIMING
  {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_split_clk, sizeof h);
    static const char *name[14] = {"pack+clear", "insert ref", "look up S1+S2", "candidates", "next pointers", "thinning (serial)", "clear bloom", "anchor table", "positions", "chain DP", "copy nxt",
                                   "chain extraction", "walk", "rest of job"};
    double tot = 0;
    for (int k = 0; k < 14; ++k) tot += (double)h[k];
    fprintf(stderr, "[split timing] %lld jobs, cycles per job (thread 0 of the CTA):", (long long)(4 * n));
    for (int k = 0; k < 14; ++k) fprintf(stderr, "  %s %.0f (%.1f%%)", name[k], (double)h[k] / (4.0 * n), 100.0 * (double)h[k] / tot);
    fprintf(stderr, "  | total %.0f\n", tot / (4.0 * n));
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_split_clk, z, sizeof z);
  }
#endif
  // sizes of the outputs, then the windows themselves
  for (int k = 0; k < 4; ++k) CU(cudaMemcpyAsync(&ctx->h_totals[k], b.base[k].as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&ctx->h_totals[4], b.misc.as<int32_t>() + 1, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int k = 0; k < 4; ++k) tot[k] = ctx->h_totals[k];
  if ((int32_t)(ctx->h_totals[4] & 0xffffffff)) return ctx->fail(ELECTOR_ECAPACITY, "a triplet has more windows than its share of the record buffer");
  for (int k = 0; k < 3; ++k) { CU(b.w_off[k].reserve((size_t)(tot[0] + 1) * 8)); CU(b.w_let[k].reserve((size_t)tot[k + 1] + 16)); }
  CU(b.rf.reserve((size_t)(n + 1) * 8));
  split_emit_kernel<<<(unsigned)n, 128, 0, st>>>(n, a.let[0], a.let[1], a.let[2], a.off[0], a.off[1], a.off[2], b.status.as<int32_t>(), b.kidx.as<int32_t>(), a.wins, a.win_off,
                                                b.base[0].as<int64_t>(), b.base[1].as<int64_t>(), b.base[2].as<int64_t>(), b.base[3].as<int64_t>(), b.w_off[0].as<int64_t>(),
                                                b.w_off[1].as<int64_t>(), b.w_off[2].as<int64_t>(), b.w_let[0].as<uint8_t>(), b.w_let[1].as<uint8_t>(), b.w_let[2].as<uint8_t>(),
                                                b.rf.as<int64_t>());
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_split1, st));
  ctx->last_launches += 10;
  return ELECTOR_OK;
}
}  // namespace

extern "C" {

int elector_split_run(elector_ctx *ctx, int64_t n, const char *ref, const int64_t *ref_off, const char *unc, const int64_t *unc_off, const char *cor,
                      const int64_t *cor_off, const int32_t *header_len, double threshold, int32_t *status, int32_t *k_used, int64_t *read_first,
                      int64_t win_cap, int64_t *w_ref_off, int64_t *w_unc_off, int64_t *w_cor_off, char *w_ref, int64_t w_ref_cap, char *w_unc,
                      int64_t w_unc_cap, char *w_cor, int64_t w_cor_cap, int64_t *n_windows) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n < 0 || (n > 0 && (!ref || !unc || !cor || !ref_off || !unc_off || !cor_off || !header_len || !read_first || !w_ref_off || !w_unc_off || !w_cor_off || !w_ref ||
                          !w_unc || !w_cor || !n_windows)))
    return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_windows) *n_windows = 0;
  if (n == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->last_launches = 0;
  int64_t tot[4];
  const int rc = split_device(ctx, n, ref, ref_off, unc, unc_off, cor, cor_off, header_len, threshold, tot);
  if (rc != ELECTOR_OK) return rc;
  if (tot[0] > win_cap || tot[1] > w_ref_cap || tot[2] > w_unc_cap || tot[3] > w_cor_cap)
    return ctx->fail(ELECTOR_ECAPACITY, "output buffers too small: %lld windows, %lld / %lld / %lld letters", (long long)tot[0], (long long)tot[1], (long long)tot[2], (long long)tot[3]);
  SplitBufs &b = *split_bufs(ctx);
  cudaStream_t st = ctx->stream;
  int64_t *w_off_h[3] = {w_ref_off, w_unc_off, w_cor_off};
  char *w_let_h[3] = {w_ref, w_unc, w_cor};
  for (int k = 0; k < 3; ++k) {
    CU(cudaMemcpyAsync(w_off_h[k], b.w_off[k].p, (size_t)(tot[0] + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(w_let_h[k], b.w_let[k].p, (size_t)tot[k + 1], cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(read_first, b.rf.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (status) CU(cudaMemcpyAsync(status, b.status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> kidx;
  if (k_used) { kidx.resize((size_t)n); CU(cudaMemcpyAsync(kidx.data(), b.kidx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st)); }
  CU(cudaStreamSynchronize(st));
  if (k_used) for (int64_t t = 0; t < n; ++t) k_used[t] = 15 - 2 * kidx[(size_t)t];
  *n_windows = tot[0];
  cudaEventElapsedTime(&ctx->last_ms, ctx->ev_split0, ctx->ev_split1);
  return ELECTOR_OK;
}

int elector_reads_run(elector_ctx *ctx, int64_t n, const char *ref, const int64_t *ref_off, const char *unc, const int64_t *unc_off, const char *cor,
                      const int64_t *cor_off, const int32_t *header_len, double threshold, int32_t *status, int32_t *k_used, int64_t *read_first,
                      int64_t *n_windows, int64_t *counters_out, int64_t *sums_out, char *m_ref, char *m_cor, char *m_unc, int64_t m_cap, int64_t *m_off,
                      int32_t *m_len) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n < 0 || (n > 0 && (!ref || !unc || !cor || !ref_off || !unc_off || !cor_off || !header_len || !counters_out))) return ctx->fail(ELECTOR_EINVAL, "null argument");
  if ((m_ref || m_cor || m_unc) && (!m_ref || !m_cor || !m_unc || !m_off || !m_len)) return ctx->fail(ELECTOR_EINVAL, "merged rows need all of m_ref, m_cor, m_unc, m_off, m_len");
  if (n_windows) *n_windows = 0;
  if (sums_out) memset(sums_out, 0, ELECTOR_TALLY_K * sizeof(int64_t));
  if (n == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->trace = getenv("ELECTOR_TRACE") != nullptr;
  ctx->last_ms = ctx->last_ms_phase1 = 0.f;
  ctx->last_launches = 0;
  int64_t tot[4];
  int rc = split_device(ctx, n, ref, ref_off, unc, unc_off, cor, cor_off, header_len, threshold, tot);
  if (rc != ELECTOR_OK) return rc;
  SplitBufs &b = *split_bufs(ctx);
  cudaStream_t st = ctx->stream;
  const int64_t nw = tot[0];
  std::vector<int64_t> rf((size_t)n + 1);
  CU(cudaMemcpyAsync(rf.data(), b.rf.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
  // the windows go to the alignment where they are: letters and offsets of the three kinds (the alignment takes ref, cor, unc)
  const int64_t rows_cap = (3 * (tot[1] + tot[2] + tot[3] + 3 * nw) + 31) & ~(int64_t)15;
  CU(ctx->d_rows.reserve(rows_cap)); CU(ctx->d_rowoff.reserve(nw * 8)); CU(ctx->d_stride.reserve(nw * 4));
  CU(ctx->d_nring.reserve(nw * 4)); CU(ctx->d_s1.reserve(nw * 4)); CU(ctx->d_s2.reserve(nw * 4)); CU(ctx->d_cells.reserve(nw * 8));
  CU(ctx->d_tally_out.reserve(n * ELECTOR_TALLY_K * 8)); CU(ctx->d_sums.reserve(ELECTOR_TALLY_K * 8));
  for (int attempt = 0;; ++attempt) {
    rc = run_device(ctx, nw, b.w_let[0].as<char>(), b.w_off[0].as<int64_t>(), b.w_let[2].as<char>(), b.w_off[2].as<int64_t>(), b.w_let[1].as<char>(),
                    b.w_off[1].as<int64_t>(), nullptr, nullptr, ctx->d_rows.as<char>(), rows_cap, ctx->d_rowoff.as<int64_t>(), ctx->d_stride.as<int32_t>(),
                    ctx->d_nring.as<int32_t>(), ctx->d_s1.as<int32_t>(), ctx->d_s2.as<int32_t>(), ctx->d_cells.as<int64_t>(), ctx->d_ctrl.as<unsigned long long>(),
                    ctx->d_ctrl.as<int32_t>() + 2);
    if (rc != ELECTOR_OK) return rc;
    CU(cudaMemcpyAsync(&ctx->h_totals[4], ctx->d_ctrl.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    rc = check_tables(ctx);
    if (rc == 1 && attempt < 3) continue;
    if (rc != ELECTOR_OK) return rc == 1 ? ctx->fail(ELECTOR_ECUDA, "scratch pools keep overflowing") : rc;
    break;
  }
  add_kernel_ms(ctx);
  if ((int32_t)(ctx->h_totals[5] & 0xffffffff)) return ctx->fail(ELECTOR_ECAPACITY, "row buffer too small (internal bound)");
  // Donatello + tally: one read per triplet (the records of a triplet share its header, Master_Splitter.cpp:283-285)
  CU(cudaMemsetAsync(ctx->d_sums.p, 0, ELECTOR_TALLY_K * 8, st));
  CU(cudaEventRecord(ctx->ev_mt0, st));
  rc = merge_device(ctx, n, rf.data(), nw, ctx->d_rows.as<uint8_t>(), rows_cap, ctx->d_rowoff.as<int64_t>(), ctx->d_stride.as<int32_t>(), ctx->d_nring.as<int32_t>());
  if (rc == ELECTOR_OK)
    rc = tally_device(ctx, n, ctx->d_mref.as<uint8_t>(), ctx->d_mcor.as<uint8_t>(), ctx->d_munc.as<uint8_t>(), ctx->d_moff.as<int64_t>(), ctx->d_mlen.as<int32_t>(),
                      ctx->d_tally_out.as<int64_t>(), ctx->merged_cap);
  if (rc != ELECTOR_OK) return rc;
  tally_sum_kernel<<<std::min<int>(64, (int)((n + 7) / 8)), 256, 0, st>>>(n, ctx->d_tally_out.as<int64_t>(), ctx->d_sums.as<unsigned long long>());
  CU(cudaGetLastError());
  ++ctx->last_launches;
  CU(cudaEventRecord(ctx->ev_mt1, st));
  CU(cudaMemcpyAsync(counters_out, ctx->d_tally_out.p, (size_t)n * ELECTOR_TALLY_K * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ctx->h_sums, ctx->d_sums.p, ELECTOR_TALLY_K * 8, cudaMemcpyDeviceToHost, st));
  if (status) CU(cudaMemcpyAsync(status, b.status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> kidx;
  if (k_used) { kidx.resize((size_t)n); CU(cudaMemcpyAsync(kidx.data(), b.kidx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st)); }
  std::vector<int64_t> moff;
  if (m_ref) {
    moff.resize((size_t)n + 1);
    CU(cudaMemcpyAsync(moff.data(), ctx->d_moff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(m_len, ctx->d_mlen.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  }
  rc = check_scan_overflow(ctx, n);   // synchronises the stream
  if (rc != ELECTOR_OK) return rc;
  float ms = 0.f;
  ctx->last_ms_split = ctx->last_ms_tally = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_split0, ctx->ev_split1) == cudaSuccess) { ctx->last_ms += ms; ctx->last_ms_split = ms; }
  if (cudaEventElapsedTime(&ms, ctx->ev_mt0, ctx->ev_mt1) == cudaSuccess) { ctx->last_ms += ms; ctx->last_ms_tally = ms; }
  if (m_ref) {
    if (moff[(size_t)n] > m_cap) return ctx->fail(ELECTOR_ECAPACITY, "merged rows need %lld bytes per buffer, capacity %lld", (long long)moff[(size_t)n], (long long)m_cap);
    memcpy(m_off, moff.data(), (size_t)n * 8);
    CU(cudaMemcpyAsync(m_ref, ctx->d_mref.p, (size_t)moff[(size_t)n], cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(m_cor, ctx->d_mcor.p, (size_t)moff[(size_t)n], cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(m_unc, ctx->d_munc.p, (size_t)moff[(size_t)n], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  if (k_used) for (int64_t t = 0; t < n; ++t) k_used[t] = 15 - 2 * kidx[(size_t)t];
  if (read_first) memcpy(read_first, rf.data(), (size_t)(n + 1) * 8);
  if (sums_out) memcpy(sums_out, ctx->h_sums, ELECTOR_TALLY_K * sizeof(int64_t));
  if (n_windows) *n_windows = nw;
  return ELECTOR_OK;
}

}  // extern "C"

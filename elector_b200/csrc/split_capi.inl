// split_capi.inl -- C-ABI entry points of the window cutting (included by capi.cu)
namespace {
struct SplitBufs {
  DevBuf let[3], off[3], hl, st_in, pool, wins, win_off, job_n, job_l, status, kidx, nrec, nl[3], base[4], w_off[3], w_let[3], rf, misc;
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
  for (DevBuf *d : {&b->let[0], &b->let[1], &b->let[2], &b->off[0], &b->off[1], &b->off[2], &b->hl, &b->st_in, &b->pool, &b->wins, &b->win_off, &b->job_n, &b->job_l,
                    &b->status, &b->kidx, &b->nrec, &b->nl[0], &b->nl[1], &b->nl[2], &b->base[0], &b->base[1], &b->base[2], &b->base[3], &b->w_off[0], &b->w_off[1],
                    &b->w_off[2], &b->w_let[0], &b->w_let[1], &b->w_let[2], &b->rf, &b->misc})
    d->release();
  delete b;
  ctx->split_state = nullptr;
}

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
  if (n > 0x1fffffff) return ctx->fail(ELECTOR_EINVAL, "too many triplets in one call");
  CU(cudaSetDevice(ctx->device));
  SplitBufs &b = *split_bufs(ctx);
  cudaStream_t st = ctx->stream;
  const char *h_let[3] = {ref, unc, cor};
  const int64_t *h_off[3] = {ref_off, unc_off, cor_off};
  // the host decides what main() decides before best_split (:412-413,:425-432): a corrected read shorter than the threshold share
  // of its reference is not cut; the longest reference read sizes the tables
  std::vector<int32_t> st_in((size_t)n);
  std::vector<int64_t> win_off((size_t)n + 1);
  int64_t longest = 0;
  win_off[0] = 0;
  for (int64_t t = 0; t < n; ++t) {
    const int64_t lr = ref_off[t + 1] - ref_off[t], lb = cor_off[t + 1] - cor_off[t];
    if (lr <= 0) return ctx->fail(ELECTOR_EINVAL, "triplet %lld has an empty reference read", (long long)t);
    if (lr > 0x3fffffff) return ctx->fail(ELECTOR_ETOOLARGE, "triplet %lld: read too long", (long long)t);
    st_in[(size_t)t] = ((double)lb / (double)lr >= threshold) ? 0 : 1;
    longest = std::max(longest, lr);
    win_off[(size_t)t + 1] = win_off[(size_t)t] + 4 * (lr / 16 + 16);
  }
  if (win_off[(size_t)n] / 4 > win_cap) return ctx->fail(ELECTOR_ECAPACITY, "win_cap %lld too small (%lld needed)", (long long)win_cap, (long long)(win_off[(size_t)n] / 4));
  uint32_t max_slots = 64;
  while (max_slots < 2 * (uint64_t)longest + 2) max_slots <<= 1;
  const int32_t max_anchors = (int32_t)(longest / 8 + 16);
  const uint32_t cand_words = (uint32_t)(longest / 32 + 2);
  const uint64_t cta_words = 2 * ((split_scratch_words(max_slots, cand_words, max_anchors) + 31) & ~31ull);
  int grid = (int)std::min<int64_t>(4 * n, (int64_t)ctx->sm_count * 4);
  while (grid > 1 && cta_words * 4 * (uint64_t)grid > ((uint64_t)16 << 30)) grid = (grid + 1) / 2;
  for (int k = 0; k < 3; ++k) {
    CU(b.let[k].reserve((size_t)(h_off[k][n] - h_off[k][0]) + 16)); CU(b.off[k].reserve((size_t)(n + 1) * 8));
    CU(cudaMemcpyAsync(b.let[k].p, h_let[k] + h_off[k][0], (size_t)(h_off[k][n] - h_off[k][0]), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b.off[k].p, h_off[k], (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  }
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
  a.pool = b.pool.as<uint32_t>(); a.cta_words = cta_words; a.max_slots = max_slots; a.max_anchors = max_anchors; a.cand_words = cand_words;
  a.wins = b.wins.as<SplitWin>(); a.win_off = b.win_off.as<int64_t>();
  a.job_n = b.job_n.as<int32_t>(); a.job_largest = b.job_l.as<uint32_t>();
  a.counter = b.misc.as<int32_t>();
  CU(cudaEventRecord(ctx->ev0, st));
  split_jobs_kernel<<<grid, 256, 0, st>>>(a);
  split_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a.status_in, a.job_n, a.job_largest, a.wins, a.win_off, b.status.as<int32_t>(), b.kidx.as<int32_t>(),
                                                                  b.nrec.as<int64_t>(), b.nl[0].as<int64_t>(), b.nl[1].as<int64_t>(), b.nl[2].as<int64_t>(), b.misc.as<int32_t>() + 1);
  scan_offsets_kernel<<<1, 1024, 0, st>>>(n, b.nrec.as<int64_t>(), b.base[0].as<int64_t>());
  for (int k = 0; k < 3; ++k) scan_offsets_kernel<<<1, 1024, 0, st>>>(n, b.nl[k].as<int64_t>(), b.base[k + 1].as<int64_t>());
  CU(cudaGetLastError());
  // sizes of the outputs, then the windows themselves
  int64_t tot[4];
  for (int k = 0; k < 4; ++k) CU(cudaMemcpyAsync(&ctx->h_totals[k], b.base[k].as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
  int32_t err = 0;
  CU(cudaMemcpyAsync(&ctx->h_totals[4], b.misc.as<int32_t>() + 1, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int k = 0; k < 4; ++k) tot[k] = ctx->h_totals[k];
  err = (int32_t)(ctx->h_totals[4] & 0xffffffff);
  if (err) return ctx->fail(ELECTOR_ECAPACITY, "a triplet has more windows than its share of the record buffer");
  if (tot[0] > win_cap || tot[1] > w_ref_cap || tot[2] > w_unc_cap || tot[3] > w_cor_cap)
    return ctx->fail(ELECTOR_ECAPACITY, "output buffers too small: %lld windows, %lld / %lld / %lld letters", (long long)tot[0], (long long)tot[1], (long long)tot[2], (long long)tot[3]);
  for (int k = 0; k < 3; ++k) { CU(b.w_off[k].reserve((size_t)(tot[0] + 1) * 8)); CU(b.w_let[k].reserve((size_t)tot[k + 1] + 16)); }
  CU(b.rf.reserve((size_t)(n + 1) * 8));
  split_emit_kernel<<<(unsigned)n, 128, 0, st>>>(n, a.let[0], a.let[1], a.let[2], a.off[0], a.off[1], a.off[2], b.status.as<int32_t>(), b.kidx.as<int32_t>(), a.wins, a.win_off,
                                                b.base[0].as<int64_t>(), b.base[1].as<int64_t>(), b.base[2].as<int64_t>(), b.base[3].as<int64_t>(), b.w_off[0].as<int64_t>(),
                                                b.w_off[1].as<int64_t>(), b.w_off[2].as<int64_t>(), b.w_let[0].as<uint8_t>(), b.w_let[1].as<uint8_t>(), b.w_let[2].as<uint8_t>(),
                                                b.rf.as<int64_t>());
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev1, st));
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
  ctx->last_launches = 8;
  cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
  return ELECTOR_OK;
}

}  // extern "C"

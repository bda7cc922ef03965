// tally_capi.inl -- C-ABI entry points of the merge + tally (included by capi.cu)
namespace {

// scan + count on device-resident merged rows; counters land in d_counters
int tally_device(elector_ctx *ctx, int64_t n_reads, const uint8_t *dR, const uint8_t *dC, const uint8_t *dU,
                 const int64_t *d_off, const int32_t *d_len, int64_t *d_counters, int64_t total_bytes) {
  if (n_reads == 0) return ELECTOR_OK;
  // dot bitmasks: read r's words start at (off[r] >> 5) + r in each of the three planes
  const int64_t plane_words = (total_bytes >> 5) + n_reads + 2;
  CU(ctx->d_tally_scan.reserve((size_t)plane_words * 3 * sizeof(uint32_t)));
  CU(ctx->d_stretch.reserve((size_t)n_reads * ELECTOR_STRETCH_K * sizeof(int32_t)));
  ctx->stretch_reads = n_reads;
  tally_read_kernel<<<(unsigned)n_reads, 128, 0, ctx->stream>>>(n_reads, dR, dC, dU, d_off, d_len, ctx->d_tally_scan.as<uint32_t>(), plane_words,
                                                               d_counters, ctx->d_stretch.as<int32_t>(), ctx->d_ctrl.as<int32_t>() + 3, ctx->d_ctrl.as<int32_t>() + kAbortWord);
  CU(cudaGetLastError());
  ctx->last_launches += 1;
#ifdef ELECTOR_TALLY_TIMING
  {
    unsigned long long h[4];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(h, g_tally_clk, sizeof h);
    fprintf(stderr, "[tally timing] clocks per read: A %.0f  B %.0f  C %.0f  reduce+store %.0f (cumulative over %lld reads)\n", (double)h[0] / n_reads, (double)h[1] / n_reads,
            (double)h[2] / n_reads, (double)h[3] / n_reads, (long long)n_reads);
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_tally_clk, z, sizeof z);
  }
#endif
  return ELECTOR_OK;
}

// offsets + merge on device-resident window rows; merged rows land in ctx->d_m{ref,cor,unc}
int merge_device(elector_ctx *ctx, int64_t n_reads, const int64_t *h_read_first, int64_t n_windows, const uint8_t *d_rows,
                 int64_t rows_bytes, const int64_t *d_row_off, const int32_t *d_row_stride, const int32_t *d_nring) {
  if (h_read_first[0] != 0 || h_read_first[n_reads] != n_windows) return ctx->fail(ELECTOR_EINVAL, "read_first must span 0..n_windows");
  const int64_t cap = rows_bytes / 3 + 16 * n_reads + 16;
  CU(ctx->d_readfirst.reserve((n_reads + 1) * 8));
  CU(ctx->d_mtot.reserve((n_reads + 1) * 8));
  CU(ctx->d_moff.reserve((n_reads + 1) * 8));
  CU(ctx->d_mlen.reserve((n_reads + 1) * 4));
  CU(ctx->d_mref.reserve(cap)); CU(ctx->d_mcor.reserve(cap)); CU(ctx->d_munc.reserve(cap));
  CU(cudaMemcpyAsync(ctx->d_readfirst.p, h_read_first, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  read_totals_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, ctx->stream>>>(n_reads, ctx->d_readfirst.as<int64_t>(), d_nring, ctx->d_mtot.as<int64_t>(), ctx->d_ctrl.as<int32_t>() + kAbortWord);
  scan_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_reads, ctx->d_mtot.as<int64_t>(), ctx->d_moff.as<int64_t>());
  CU(ctx->d_wdst.reserve((size_t)n_windows * 8));
  merge_plan_kernel<<<(unsigned)((n_reads + 3) / 4), 128, 0, ctx->stream>>>(n_reads, ctx->d_readfirst.as<int64_t>(), d_rows, d_row_off, d_row_stride, d_nring,
                                                                            ctx->d_moff.as<int64_t>(), ctx->d_wdst.as<int64_t>(), ctx->d_mlen.as<int32_t>(), ctx->d_ctrl.as<int32_t>() + kAbortWord);
  merge_copy_kernel<<<(unsigned)std::min<int64_t>((n_windows + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(   // a warp takes 32 windows at a time
      n_windows, d_rows, d_row_off, d_row_stride, d_nring, ctx->d_wdst.as<int64_t>(), ctx->d_mref.as<uint8_t>(), ctx->d_mcor.as<uint8_t>(),
      ctx->d_munc.as<uint8_t>(), ctx->d_ctrl.as<int32_t>() + kAbortWord);
  CU(cudaGetLastError());
  ctx->last_launches += 4;
  ctx->merged_cap = cap;
  return ELECTOR_OK;
}

// d_ctrl[3]: set by tally_scan_kernel when a read has more border gap stretches than ReadScan holds
int check_scan_overflow(elector_ctx *ctx, int64_t n_reads) {
  (void)n_reads;
  int32_t flag = 0;
  CU(cudaMemcpyAsync(&flag, ctx->d_ctrl.as<int32_t>() + 3, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (flag) {
    cudaMemsetAsync(ctx->d_ctrl.as<int32_t>() + 3, 0, sizeof flag, ctx->stream);
    return ctx->fail(ELECTOR_EUNSUPPORTED, "a read has more than %d gap stretches at its borders", kMaxStretchKeys);
  }
  return ELECTOR_OK;
}

}  // namespace

extern "C" {

int elector_tally_run(elector_ctx *ctx, int64_t n_reads, const char *row_ref, const char *row_cor, const char *row_unc,
                      const int64_t *row_off, int64_t *counters_out) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_reads < 0 || (n_reads > 0 && (!row_ref || !row_cor || !row_unc || !row_off || !counters_out))) return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_reads == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  const int64_t bytes = row_off[n_reads];
  CU(ctx->d_mref.reserve(bytes + 16)); CU(ctx->d_mcor.reserve(bytes + 16)); CU(ctx->d_munc.reserve(bytes + 16));
  CU(ctx->d_moff.reserve((n_reads + 1) * 8));
  CU(ctx->d_tally_out.reserve(n_reads * ELECTOR_TALLY_K * 8));
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(ctx->d_mref.p, row_ref, bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_mcor.p, row_cor, bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_munc.p, row_unc, bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_moff.p, row_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  ctx->last_launches = 0;
  CU(cudaEventRecord(ctx->ev0, st));
  int rc = tally_device(ctx, n_reads, ctx->d_mref.as<uint8_t>(), ctx->d_mcor.as<uint8_t>(), ctx->d_munc.as<uint8_t>(),
                        ctx->d_moff.as<int64_t>(), nullptr, ctx->d_tally_out.as<int64_t>(), bytes);
  if (rc != ELECTOR_OK) return rc;
  CU(cudaEventRecord(ctx->ev1, st));
  CU(cudaMemcpyAsync(counters_out, ctx->d_tally_out.p, n_reads * ELECTOR_TALLY_K * 8, cudaMemcpyDeviceToHost, st));
  rc = check_scan_overflow(ctx, n_reads);
  cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
  return rc;
}

int elector_last_stretches(elector_ctx *ctx, int64_t n_reads, int32_t *stretches_out) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_reads < 0 || (n_reads > 0 && !stretches_out)) return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_reads != ctx->stretch_reads) return ctx->fail(ELECTOR_EINVAL, "the last tally on this context had %lld reads, not %lld", (long long)ctx->stretch_reads, (long long)n_reads);
  if (n_reads == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(stretches_out, ctx->d_stretch.p, (size_t)n_reads * ELECTOR_STRETCH_K * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return ELECTOR_OK;
}

int elector_merge_run(elector_ctx *ctx, int64_t n_reads, const int64_t *read_first, int64_t n_windows, const char *rows,
                      int64_t rows_bytes, const int64_t *row_off, const int32_t *row_stride, const int32_t *nring,
                      char *m_ref, char *m_cor, char *m_unc, int64_t m_cap, int64_t *m_off, int32_t *m_len) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_reads < 0 || n_windows < 0 || (n_reads > 0 && (!read_first || !rows || !row_off || !row_stride || !nring || !m_ref || !m_cor || !m_unc || !m_off || !m_len)))
    return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_reads == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CU(ctx->d_rows.reserve(rows_bytes)); CU(ctx->d_rowoff.reserve(n_windows * 8)); CU(ctx->d_stride.reserve(n_windows * 4)); CU(ctx->d_nring.reserve(n_windows * 4));
  CU(cudaMemcpyAsync(ctx->d_rows.p, rows, rows_bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_rowoff.p, row_off, n_windows * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_stride.p, row_stride, n_windows * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_nring.p, nring, n_windows * 4, cudaMemcpyHostToDevice, st));
  ctx->last_launches = 0;
  int rc = merge_device(ctx, n_reads, read_first, n_windows, ctx->d_rows.as<uint8_t>(), rows_bytes, ctx->d_rowoff.as<int64_t>(),
                        ctx->d_stride.as<int32_t>(), ctx->d_nring.as<int32_t>());
  if (rc != ELECTOR_OK) return rc;
  std::vector<int64_t> off(n_reads + 1);
  CU(cudaMemcpyAsync(off.data(), ctx->d_moff.p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(m_len, ctx->d_mlen.p, n_reads * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (off[n_reads] > m_cap) return ctx->fail(ELECTOR_ECAPACITY, "merged rows need %lld bytes per buffer, capacity %lld", (long long)off[n_reads], (long long)m_cap);
  memcpy(m_off, off.data(), n_reads * 8);
  CU(cudaMemcpyAsync(m_ref, ctx->d_mref.p, off[n_reads], cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(m_cor, ctx->d_mcor.p, off[n_reads], cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(m_unc, ctx->d_munc.p, off[n_reads], cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ELECTOR_OK;
}

int elector_merge_tally_device(elector_ctx *ctx, int64_t n_reads, const int64_t *h_read_first, int64_t n_windows,
                               const char *d_rows, int64_t rows_bytes, const int64_t *d_row_off, const int32_t *d_row_stride,
                               const int32_t *d_nring, int64_t *d_counters_out) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_reads < 0 || (n_reads > 0 && (!h_read_first || !d_rows || !d_row_off || !d_row_stride || !d_nring || !d_counters_out)))
    return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_reads == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->last_launches = 0;
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = merge_device(ctx, n_reads, h_read_first, n_windows, (const uint8_t *)d_rows, rows_bytes, d_row_off, d_row_stride, d_nring);
  if (rc != ELECTOR_OK) return rc;
  rc = tally_device(ctx, n_reads, ctx->d_mref.as<uint8_t>(), ctx->d_mcor.as<uint8_t>(), ctx->d_munc.as<uint8_t>(),
                    ctx->d_moff.as<int64_t>(), ctx->d_mlen.as<int32_t>(), d_counters_out, ctx->merged_cap);
  if (rc != ELECTOR_OK) return rc;
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  rc = check_scan_overflow(ctx, n_reads);
  cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
  return rc;
}

int elector_tally_sum_device(elector_ctx *ctx, int64_t n_reads, const int64_t *d_counters, int64_t *d_sums) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_reads < 0 || (n_reads > 0 && (!d_counters || !d_sums))) return ctx->fail(ELECTOR_EINVAL, "null argument");
  if (n_reads == 0) return ELECTOR_OK;
  CU(cudaSetDevice(ctx->device));
  tally_sum_kernel<<<std::min<int>(64, (int)((n_reads + 7) / 8)), 256, 0, ctx->stream>>>(n_reads, d_counters, reinterpret_cast<unsigned long long *>(d_sums));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return ELECTOR_OK;
}

}  // extern "C"

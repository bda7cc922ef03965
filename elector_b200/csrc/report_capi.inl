// report_capi.inl -- C-ABI entry points of the report (included by capi.cu)
namespace {
int report_store(const elector::ReportResult &res, const char *out_dir, const char *soft, const char *size_file_name, elector_report_summary *summary, char *log_text,
                 int64_t log_cap, char *stdout_text, int64_t stdout_cap) {
  if (summary) *summary = res.s;
  auto put = [](const std::string &t, char *dst, int64_t cap) { if (dst && cap > 0) { const size_t n = std::min<size_t>(t.size(), (size_t)cap - 1); memcpy(dst, t.data(), n); dst[n] = 0; } };
  put(res.log_text, log_text, log_cap);
  put(res.stdout_text, stdout_text, stdout_cap);
  if (out_dir) {
    const std::string dir(out_dir);
    const std::string metrics = dir + "/" + (soft ? std::string(soft) + "_per_read_metrics.txt" : std::string("per_read_metrics.txt"));
    const std::string sizes = dir + "/" + (size_file_name ? size_file_name : "read_size_distribution.txt");
    for (const auto &f : {std::make_pair(metrics, &res.per_read_metrics), std::make_pair(sizes, &res.size_distribution)}) {
      FILE *o = fopen(f.first.c_str(), "wb");
      if (!o) { g_init_error = "cannot write " + f.first; return ELECTOR_EIO; }
      const bool ok = fwrite(f.second->data(), 1, f.second->size(), o) == f.second->size();
      if (fclose(o) != 0 || !ok) { g_init_error = "cannot write " + f.first; return ELECTOR_EIO; }
    }
  }
  return ELECTOR_OK;
}
}  // namespace

extern "C" {

int elector_report_write(int64_t n_records, const char *headers, const int64_t *header_off, const int64_t *counters, const int32_t *stretches, const char *m_ref,
                         const char *m_cor, const int64_t *m_off, int small_reads, int wrongly_cor_reads, double size_threshold, int homopolymer_threshold, int compensated_sum,
                         const char *corrected_fasta, const char *out_dir, const char *soft, const char *size_file_name, elector_report_summary *summary,
                         char *log_text, int64_t log_cap, char *stdout_text, int64_t stdout_cap) {
  if (n_records <= 0 || !headers || !header_off || !counters || !stretches || (m_ref && !m_off) || (m_cor && !m_off)) { g_init_error = "elector_report_write: null argument"; return ELECTOR_EINVAL; }
  elector::ReportRecords in;
  in.n = n_records; in.headers = headers; in.header_off = header_off; in.counters = counters; in.stretches = stretches; in.m_ref = m_ref; in.m_cor = m_cor; in.m_off = m_off;
  elector::ReportResult res;
  const int rc = elector::report_compute(in, small_reads, wrongly_cor_reads, size_threshold, homopolymer_threshold, corrected_fasta, soft, compensated_sum != 0, res);
  if (rc != ELECTOR_OK) { g_init_error = res.error; return rc; }
  return report_store(res, out_dir, soft, size_file_name, summary, log_text, log_cap, stdout_text, stdout_cap);
}

int elector_report_run(elector_ctx *ctx, int64_t n_records, const char *headers, const int64_t *header_off, const char *m_ref, const char *m_cor, const char *m_unc,
                       const int64_t *m_off, int small_reads, int wrongly_cor_reads, double size_threshold, int homopolymer_threshold, int compensated_sum, const char *corrected_fasta,
                       const char *out_dir, const char *soft, const char *size_file_name, elector_report_summary *summary, char *log_text, int64_t log_cap,
                       char *stdout_text, int64_t stdout_cap) {
  if (!ctx) return ELECTOR_EINVAL;
  if (n_records <= 0 || !headers || !header_off || !m_ref || !m_cor || !m_unc || !m_off) return ctx->fail(ELECTOR_EINVAL, "null argument");
  std::vector<int64_t> counters((size_t)n_records * ELECTOR_TALLY_K);
  std::vector<int32_t> stretches((size_t)n_records * ELECTOR_STRETCH_K);
  int rc = elector_tally_run(ctx, n_records, m_ref, m_cor, m_unc, m_off, counters.data());
  if (rc == ELECTOR_OK) rc = elector_last_stretches(ctx, n_records, stretches.data());
  if (rc != ELECTOR_OK) return rc;
  rc = elector_report_write(n_records, headers, header_off, counters.data(), stretches.data(), m_ref, m_cor, m_off, small_reads, wrongly_cor_reads, size_threshold,
                            homopolymer_threshold, compensated_sum, corrected_fasta, out_dir, soft, size_file_name, summary, log_text, log_cap, stdout_text, stdout_cap);
  if (rc != ELECTOR_OK) return ctx->fail(rc, "%s", g_init_error.c_str());
  return ELECTOR_OK;
}

}  // extern "C"

// ---- file preparation in front of the splitter ----
extern "C" {

int elector_sort_fasta(const char *in_path, const char *out_path, int64_t *n_records) {
  if (!in_path || !out_path) { g_init_error = "elector_sort_fasta: null argument"; return ELECTOR_EINVAL; }
  const int64_t n = elector::prep_sort_fasta(in_path, out_path);
  if (n < 0) { g_init_error = std::string("cannot read ") + in_path + " or write " + out_path; return ELECTOR_EIO; }
  if (n_records) *n_records = n;
  return ELECTOR_OK;
}

int elector_duplicate_reads(const char *sorted_ref, const char *sorted_unc, const char *sorted_cor, const char *new_ref, const char *new_unc, int64_t *n_triplets) {
  if (!sorted_ref || !sorted_unc || !sorted_cor || !new_ref || !new_unc) { g_init_error = "elector_duplicate_reads: null argument"; return ELECTOR_EINVAL; }
  const int64_t n = elector::prep_duplicate(sorted_ref, sorted_unc, sorted_cor, new_ref, new_unc);
  if (n == -2) { g_init_error = std::string(sorted_ref) + " does not start with a header line"; return ELECTOR_EINVAL; }
  if (n < 0) { g_init_error = "cannot read the sorted files or write the duplicated ones"; return ELECTOR_EIO; }
  if (n_triplets) *n_triplets = n;
  return ELECTOR_OK;
}

}  // extern "C"

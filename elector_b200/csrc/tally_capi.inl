// tally_capi.inl -- C-ABI entry points of the merge + tally (included by capi.cu)
extern "C" {
int elector_tally_run(elector_ctx *ctx, int64_t, const char *, const char *, const char *, const int64_t *, int64_t *) {
  if (!ctx) return ELECTOR_EINVAL;
  return ctx->fail(ELECTOR_EUNSUPPORTED, "tally kernels not built yet");
}
int elector_merge_tally_device(elector_ctx *ctx, int64_t, const int64_t *, int64_t, const char *, const int64_t *,
                               const int32_t *, const int32_t *, int64_t *) {
  if (!ctx) return ELECTOR_EINVAL;
  return ctx->fail(ELECTOR_EUNSUPPORTED, "tally kernels not built yet");
}
}

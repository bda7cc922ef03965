/* C host of the pipelined entry point (elector_pipeline_run) on raw CSR arrays written by tools/dump_csr.py:
 * the call a C caller makes, timed with the host clock; also the process ncu profiles (no Python inside).
 *   pipe_driver PREFIX [calls]                                                                      */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda_runtime_api.h>

#include "../include/elector_poa.h"

static void *slurp(const char *pre, const char *name, size_t *bytes, int pinned) {
  char path[1024];
  snprintf(path, sizeof path, "%s.%s.bin", pre, name);
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(1); }
  fseek(f, 0, SEEK_END);
  *bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void *p = NULL;
  if (pinned) { if (cudaMallocHost(&p, *bytes + 16) != cudaSuccess) exit(2); } else p = malloc(*bytes + 16);
  if (fread(p, 1, *bytes, f) != *bytes) exit(3);
  fclose(f);
  return p;
}
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec / 1e6; }

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s PREFIX [calls]\n", argv[0]); return 2; }
  const int calls = argc > 2 ? atoi(argv[2]) : 3;
  elector_ctx *ctx = NULL;
  if (elector_poa_init(0, NULL, &ctx) != ELECTOR_OK) { fprintf(stderr, "init: %s\n", elector_last_error(NULL)); return 1; }
  size_t b;
  char *ref = slurp(argv[1], "ref", &b, 1), *cor = slurp(argv[1], "cor", &b, 1), *unc = slurp(argv[1], "unc", &b, 1);
  int64_t *ro = slurp(argv[1], "ref_off", &b, 1);
  const int64_t n = (int64_t)(b / 8) - 1;
  int64_t *co = slurp(argv[1], "cor_off", &b, 1), *uo = slurp(argv[1], "unc_off", &b, 1);
  int64_t *rf = slurp(argv[1], "read_first", &b, 0);
  const int64_t n_reads = (int64_t)(b / 8) - 1;
  const int64_t bound = elector_poa_rows_bound(n, ro, co, uo);
  char *rows; int64_t *row_off, *counters, sums[ELECTOR_TALLY_K]; int32_t *stride, *nring;
  if (cudaMallocHost((void **)&rows, bound) != cudaSuccess || cudaMallocHost((void **)&row_off, n * 8) != cudaSuccess ||
      cudaMallocHost((void **)&stride, n * 4) != cudaSuccess || cudaMallocHost((void **)&nring, n * 4) != cudaSuccess ||
      cudaMallocHost((void **)&counters, n_reads * ELECTOR_TALLY_K * 8) != cudaSuccess) return 2;
  for (int k = 0; k < calls; ++k) {
    const double t0 = now_ms();
    const int rc = elector_pipeline_run(ctx, n, ref, ro, cor, co, unc, uo, n_reads, rf, rows, bound, row_off, stride, nring, NULL, NULL, NULL, counters, sums);
    const double t1 = now_ms();
    if (rc != ELECTOR_OK) { fprintf(stderr, "run: %s\n", elector_last_error(ctx)); return 1; }
    float ms = 0; int launches = 0;
    elector_last_kernel_ms(ctx, &ms, &launches);
    printf("call %d: %lld windows, %lld reads, %.2f ms (%.0f triplets/s), %d launches, %.2f ms of kernels, TP %lld\n", k, (long long)n, (long long)n_reads,
           t1 - t0, n_reads / ((t1 - t0) * 1e-3), launches, ms, (long long)sums[0]);
  }
  elector_poa_free(ctx);
  return 0;
}

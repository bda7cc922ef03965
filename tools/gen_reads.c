/* gen_reads.c -- deterministic synthetic read triplets for the BASELINE.json configs
 * (SURVEY.md 8d).  Bench / test infrastructure: writes the three 2-line-per-record FASTA
 * files that ELECTOR's alignment stage starts from (reference_sorted_duplicated.fa,
 * uncorrected_sorted_duplicated.fa, corrected_sorted.fa; elector/readAndSortFiles.py:150-191).
 *
 *   reference  i.i.d. uniform ACGT
 *   raw        reference with error events at `raw_rate` per base, ins:del:sub = I:D:S,
 *              optional homopolymer bias (x1.5 event rate on a repeated base)
 *   corrected  the same events, each kept with probability `keep`
 *   trimming / splitting of the corrected read (config 2): a trimmed read loses U(200,2000)
 *   bases at one end; a split read becomes 2-3 fragments separated by U(200,1500) missing
 *   bases, and the reference / raw records are duplicated per fragment with suffix _k.
 * PRNG: splitmix64 (same stream as oracle/synth.py).
 *
 * usage: gen_reads CONFIG N_READS FIRST_READ OUT_PREFIX      (CONFIG = 1..4)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t S;
static uint64_t nxt(void)
{
  uint64_t z;
  S += 0x9E3779B97F4A7C15ULL;
  z = S;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static double unit(void) { return (double)(nxt() >> 11) / 9007199254740992.0; }
static uint64_t below(uint64_t n) { return nxt() % n; }

typedef struct {
  uint64_t seed;
  int len_lo, len_hi, loguniform;
  double raw_rate, keep;
  int I, D, Sb, hp_bias;
  double p_trim, p_split;
} cfg_t;

static const cfg_t CFG[5] = {
  {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
  {1, 10000, 10000, 0, 0.10, 0.10, 1, 1, 1, 0, 0.0, 0.0},
  {2, 15000, 15000, 0, 0.12, 0.125, 2, 3, 3, 1, 0.20, 0.10},
  {3, 50000, 100000, 0, 0.10, 0.10, 1, 1, 1, 0, 0.0, 0.0},
  {4, 1000, 30000, 1, 0.10, 0.10, 1, 1, 1, 0, 0.0, 0.0},
};

int main(int argc, char **argv)
{
  static const char AB[4] = {'A', 'C', 'G', 'T'};
  int c, n, first, i;
  char path[4096];
  FILE *fr, *fu, *fc;
  const cfg_t *g;
  char *ref, *raw, *cor;
  if (argc < 5) { fprintf(stderr, "usage: %s CONFIG N_READS FIRST_READ OUT_PREFIX\n", argv[0]); return 2; }
  c = atoi(argv[1]); n = atoi(argv[2]); first = atoi(argv[3]);
  if (c < 1 || c > 4) return 2;
  g = &CFG[c];
  snprintf(path, sizeof path, "%s.ref.fa", argv[4]); fr = fopen(path, "w");
  snprintf(path, sizeof path, "%s.unc.fa", argv[4]); fu = fopen(path, "w");
  snprintf(path, sizeof path, "%s.cor.fa", argv[4]); fc = fopen(path, "w");
  if (!fr || !fu || !fc) return 1;
  ref = (char *)malloc(g->len_hi + 8); raw = (char *)malloc(2 * g->len_hi + 8); cor = (char *)malloc(2 * g->len_hi + 8);
  for (i = first; i < first + n; i++) {
    int L, k, nr = 0, nc = 0, tot = g->I + g->D + g->Sb;
    char prev = 0;
    /* one independent stream per read so that slices generated in parallel agree */
    S = g->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i * 0xD1B54A32D192ED03ULL;
    nxt();
    if (g->loguniform) L = (int)exp(log((double)g->len_lo) + unit() * (log((double)g->len_hi) - log((double)g->len_lo)));
    else L = g->len_lo + (g->len_hi > g->len_lo ? (int)below((uint64_t)(g->len_hi - g->len_lo + 1)) : 0);
    for (k = 0; k < L; k++) ref[k] = AB[nxt() & 3];
    for (k = 0; k < L; k++) {
      char ch = ref[k];
      double rate = g->raw_rate;
      if (g->hp_bias && ch == prev) rate *= 1.5;
      prev = ch;
      if (unit() < rate) {
        int e = (int)below((uint64_t)tot);
        int kept = unit() < g->keep;
        if (e < g->I) { char b = AB[nxt() & 3]; raw[nr++] = b; raw[nr++] = ch; if (kept) cor[nc++] = b; cor[nc++] = ch; }
        else if (e < g->I + g->D) { if (!kept) cor[nc++] = ch; }
        else { char b = AB[nxt() & 3]; raw[nr++] = b; cor[nc++] = kept ? b : ch; }
      } else { raw[nr++] = ch; cor[nc++] = ch; }
    }
    {
      double t = unit();
      int nfrag = 1, f, start[3], end[3];
      start[0] = 0; end[0] = nc;
      if (t < g->p_trim && nc > 4500) {
        int cut = 200 + (int)below(1801);
        if (nxt() & 1) start[0] = cut; else end[0] = nc - cut;
      } else if (t < g->p_trim + g->p_split && nc > 9000) {
        int pos = 0;
        nfrag = 2 + (int)(nxt() & 1);
        for (f = 0; f < nfrag; f++) {
          int gap = 200 + (int)below(1301);
          int seg = (nc - (nfrag - 1) * 1500) / nfrag;
          start[f] = pos; end[f] = (f == nfrag - 1) ? nc : pos + seg;
          pos = end[f] + gap;
          if (f == nfrag - 2 && pos > nc - 500) pos = nc - 500;
        }
      }
      for (f = 0; f < nfrag; f++) {
        fprintf(fr, ">read_%d_%d\n", i, f); fwrite(ref, 1, L, fr); fputc('\n', fr);
        fprintf(fu, ">read_%d_%d\n", i, f); fwrite(raw, 1, nr, fu); fputc('\n', fu);
        fprintf(fc, ">read_%d\n", i); fwrite(cor + start[f], 1, end[f] - start[f], fc); fputc('\n', fc);
      }
    }
  }
  fclose(fr); fclose(fu); fclose(fc);
  free(ref); free(raw); free(cor);
  return 0;
}

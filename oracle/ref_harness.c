/* TEST INFRASTRUCTURE ONLY (oracle/): a small driver of OUR OWN that calls the
 * reference's public library functions (declared in /root/reference/src/poa-graph/
 * lpo.h, seq_util.h, align_score.h) to expose what the reference `poa` binary
 * computes but never prints: the two DP scores and the x_to_y / y_to_x maps of
 * each align_lpo_po call (align_lpo_po2.c:178, score print commented out at
 * :438-439).  It is compiled by oracle/build_ref.sh against the reference's own
 * object files into oracle/_ref/ref_dump and used only to generate / check golden
 * vectors.  It performs the same call sequence as the reference:
 *   main.c:265-274      per-window loop, order ref, corrected, uncorrected
 *   buildup_lpo.c:562   buildup_pairwise_lpo = align_lpo_po + fuse_lpo
 *   main.c:364          write_lpo_bundle_as_fasta
 * and the PIR it writes is compared byte-for-byte with the real `poa` output in
 * oracle/make_golden.py, which validates the harness itself.
 *
 * usage: ref_dump MATRIX REF.fa COR.fa UNC.fa OUT.pir OUT.dump
 */
#include "lpo.h"
#include "msa_format.h"
#include "align_score.h"

static void dump_map(FILE *f, const char *tag, int n, LPOLetterRef_T *m)
{
  int i;
  fprintf(f, "%s %d", tag, n);
  for (i = 0; i < n; i++) fprintf(f, " %d", (int)m[i]);
  fputc('\n', f);
}

int main(int argc, char *argv[])
{
  ResidueScoreMatrix_T m;
  LPOSequence_T *ref = NULL, *cor = NULL, *unc = NULL;
  char *comment = NULL;
  FILE *f, *pir, *dump;
  int n, i;

  if (argc < 7) {
    fprintf(stderr, "usage: %s MATRIX REF.fa COR.fa UNC.fa OUT.pir OUT.dump\n", argv[0]);
    return 2;
  }
  black_flag_init(argv[0], "ref_dump");
  if (read_score_matrix(argv[1], &m) <= 0) return 1;
  if (!(f = fopen(argv[3], "r"))) return 1;
  n = read_fasta(f, &cor, switch_case_to_lower, &comment); fclose(f);
  if (!(f = fopen(argv[4], "r"))) return 1;
  n = read_fasta(f, &unc, switch_case_to_lower, &comment); fclose(f);
  if (!(f = fopen(argv[2], "r"))) return 1;
  n = read_fasta(f, &ref, switch_case_to_lower, &comment); fclose(f);
  if (n == 0) return 1;
  pir = fopen(argv[5], "w");
  dump = fopen(argv[6], "w");
  if (!pir || !dump) return 1;

  for (i = 0; i < n; i++) {
    LPOLetterRef_T *x2y = NULL, *y2x = NULL;
    LPOScore_T s1, s2;
    int lx, ly, lp1;

    initialize_seqs_as_lpo(1, &ref[i], &m);
    initialize_seqs_as_lpo(1, &cor[i], &m);
    initialize_seqs_as_lpo(1, &unc[i], &m);

    lx = ref[i].length; ly = cor[i].length;
    s1 = align_lpo_po(&ref[i], &cor[i], &m, &x2y, &y2x, matrix_scoring_function, 1);
    fprintf(dump, "W %d %d %d %d\n", i, lx, ly, unc[i].length);
    fprintf(dump, "S1 %d\n", (int)s1);
    dump_map(dump, "X1", lx, x2y);
    dump_map(dump, "Y1", ly, y2x);
    fuse_lpo(&ref[i], &cor[i], x2y, y2x);
    free(x2y); free(y2x); x2y = y2x = NULL;

    lp1 = ref[i].length; ly = unc[i].length;
    s2 = align_lpo_po(&ref[i], &unc[i], &m, &x2y, &y2x, matrix_scoring_function, 1);
    fprintf(dump, "S2 %d %d\n", (int)s2, lp1);
    dump_map(dump, "X2", lp1, x2y);
    dump_map(dump, "Y2", ly, y2x);
    fuse_lpo(&ref[i], &unc[i], x2y, y2x);
    free(x2y); free(y2x);
    fprintf(dump, "L %d\n", ref[i].length);

    write_lpo_bundle_as_fasta(pir, &ref[i], m.nsymbol, m.symbol, ALL_BUNDLES);
  }
  fclose(pir);
  fclose(dump);
  return 0;
}

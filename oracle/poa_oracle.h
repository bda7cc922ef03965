/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, flat arrays) of the
 * reference's POA hot path -- see poa_oracle.c for the file:line map.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this.  The product (elector_b200/) never does.
 */
#ifndef POA_ORACLE_H
#define POA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXSYM 128
#define ORA_MAXPRED 8
#define ORA_MAXSRC 3

typedef struct {
  int nsymbol;
  char symbol[ORA_MAXSYM + 1];
  int score[ORA_MAXSYM][ORA_MAXSYM];
  int gap_set[2][3];
  int trunc_gap_length, decay_gap_length, max_gap_length;
  int gap_penalty_x[64], gap_penalty_y[64];
} ora_matrix;

typedef struct {
  int n;                 /* number of nodes (letters) */
  unsigned char *letter; /* symbol index per node */
  int *npred;            /* predecessor count per node */
  int *pred;             /* [n][ORA_MAXPRED], list order preserved */
  int *src;              /* [n][ORA_MAXSRC] position in source sequence or -1 */
  int *ring_id;          /* minimum node index on the node's align ring */
  int *align_ring;       /* circular list of aligned nodes */
  int nsrc;              /* number of source sequences */
  int src_len[ORA_MAXSRC];
} ora_po;

/* result of one window (ref, cor, unc) */
typedef struct {
  int score1, score2;    /* best_score of the two align_lpo_po calls */
  int len_p1, len_p2;    /* PO lengths after fuse 1 / fuse 2 */
  int nring;             /* MSA columns */
  long long cells;       /* lr*lc + len_p1*lu  (DP inner-loop iterations) */
  char *rows;            /* 3*nring chars, row-major (caller frees via ora_free_result) */
  int *x2y1, *y2x1, *x2y2, *y2x2; /* alignment maps of both DPs */
} ora_result;

int ora_read_matrix(const char *path, ora_matrix *m); /* returns nsymbol or <=0 */
void ora_default_matrix(ora_matrix *m);               /* the shipped blosum80.mat values */

/* normalise raw FASTA letters: lower-case, limit to alphabet, index */
void ora_index_sequence(const ora_matrix *m, const char *seq, int len, unsigned char *out);

void ora_po_linear(ora_po *p, const unsigned char *codes, int len);
void ora_po_free(ora_po *p);
int ora_align(const ora_po *x, const ora_po *y, const ora_matrix *m, int *x2y, int *y2x);
void ora_fuse(ora_po *x, const ora_po *y, const int *x2y, const int *y2x);
int ora_emit(const ora_po *p, const ora_matrix *m, char **rows_out);

int ora_window(const ora_matrix *m, const char *ref, int lr, const char *cor, int lc,
               const char *unc, int lu, ora_result *res);
void ora_free_result(ora_result *res);

/* batch entry (CSR-concatenated sequences), used through ctypes by the tests and
 * by bench.py's cpu_baseline leg; nthreads>1 uses OpenMP over windows.
 * rows_out must hold 3*sum(lr+lc+lu) bytes; row_off[w] receives the offset of
 * window w's 3*nring block.  Returns 0. */
int ora_batch(const ora_matrix *m, int n, const char *ref, const long long *ref_off,
              const char *cor, const long long *cor_off, const char *unc,
              const long long *unc_off, char *rows_out, long long *row_off, int *nring,
              int *score1, int *score2, long long *cells, int nthreads);

/* `poa`-like file driver: same flags' semantics as main.c (FASTA in, PIR out) */
int ora_poa_files(const char *matrix, const char *ref_fa, const char *cor_fa,
                  const char *unc_fa, const char *pir_out, int print_perm);

/* writes scores + alignment maps in the text format of oracle/ref_harness.c */
int ora_dump_files(const char *matrix, const char *ref_fa, const char *cor_fa,
                   const char *unc_fa, const char *dump_out);

#ifdef __cplusplus
}
#endif
#endif

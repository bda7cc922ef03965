/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the ELECTOR POA hot path.
 *
 * A from-scratch restatement (flat arrays, no linked lists) of what the reference
 * `poa` program computes for one (reference, corrected, uncorrected) window.  It is
 * deliberately GENERAL (arbitrary partial orders on both sides, full 0..M+1 gap-length
 * state, arbitrary score matrix) so that it checks the specialisations the CUDA path
 * makes, instead of sharing them.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file byte-for-byte against
 * committed outputs of the compiled reference (`oracle/_ref/poa`, `ref_dump`) --
 * PIR rows, both DP scores and all four alignment maps -- see tests/golden/.
 *
 * Reference map (paths relative to /root/reference/src/poa-graph/):
 *   ora_read_matrix      seq_util.c:82-217   read_score_matrix
 *   ora_index_sequence   create_seq.c:39-43 (lower-casing), seq_util.c:253-263
 *                        limit_residues, seq_util.c:37-52 index_symbols
 *   ora_po_linear        lpo.c:11-32 lpo_init
 *   node_stats           align_lpo_po2.c:29-105 get_lpo_stats
 *   ora_align            align_lpo_po2.c:178-487 align_lpo_po (+ :108-168 traceback)
 *   reindex_fusion       lpo.c:413-463 reindex_lpo_fusion, :369-382 mark_fusion_segments
 *   ora_fuse             lpo.c:602-656 fuse_lpo_remap, :577-598 translate_lpo,
 *                        :308-320 copy_lpo_letter, :227-241 add_lpo_link,
 *                        :246-258 add_lpo_sources, :325-359 crosslink / copy ring
 *   ora_emit             lpo_format.c:337-393 xlate_lpo_to_al
 *   ora_window           main.c:265-274, buildup_lpo.c:408-589 (merge order ref<-cor, PO<-unc)
 *   fasta reader         fasta_format.c:10-66 read_fasta, create_seq.c:22-51
 *   ora_poa_files        main.c:241-287, lpo_format.c:398-426 write_lpo_bundle_as_fasta
 */
#include "poa_oracle.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MIN_SCORE (-999999)

/* ------------------------------------------------------------------ matrix */

static void build_gap_tables(ora_matrix *m)
{
  int i;
  int T = m->trunc_gap_length, D = m->decay_gap_length;
  m->max_gap_length = T + D;
  memset(m->gap_penalty_x, 0, sizeof m->gap_penalty_x);
  memset(m->gap_penalty_y, 0, sizeof m->gap_penalty_y);
  m->gap_penalty_x[0] = m->gap_set[0][0];
  m->gap_penalty_y[0] = m->gap_set[1][0];
  for (i = 1; i < T; i++) {
    m->gap_penalty_x[i] = m->gap_set[0][1];
    m->gap_penalty_y[i] = m->gap_set[1][1];
  }
  for (i = 0; i < D; i++) {
    double dx = (m->gap_set[0][1] - m->gap_set[0][2]) / ((double)(D + 1));
    double dy = (m->gap_set[1][1] - m->gap_set[1][2]) / ((double)(D + 1));
    m->gap_penalty_x[i + T] = (int)(m->gap_set[0][1] - (i + 1) * dx);
    m->gap_penalty_y[i + T] = (int)(m->gap_set[1][1] - (i + 1) * dy);
  }
  m->gap_penalty_x[T + D] = m->gap_set[0][2];
  m->gap_penalty_y[T + D] = m->gap_set[1][2];
  m->gap_penalty_x[T + D + 1] = 0;
  m->gap_penalty_y[T + D + 1] = 0;
}

int ora_read_matrix(const char *path, ora_matrix *m)
{
  char line[1024];
  int i, j, k, nsymb = 0, have_symbols = 0, isymb;
  FILE *f = fopen(path, "r");
  if (!f) return -2;
  memset(m, 0, sizeof *m);
  m->gap_set[0][0] = m->gap_set[1][0] = 12;
  m->gap_set[0][1] = m->gap_set[1][1] = 2;
  m->gap_set[0][2] = m->gap_set[1][2] = 0;
  m->trunc_gap_length = 10; /* TRUNCATE_GAP_LENGTH, seq_util.h */
  m->decay_gap_length = 5;  /* DECAY_GAP_LENGTH */
  while (fgets(line, 1023, f)) {
    if (line[0] == '#' || line[0] == '\n') continue;
    if (sscanf(line, "GAP-TRUNCATION-LENGTH=%d", &i) == 1) { m->trunc_gap_length = i; continue; }
    if (sscanf(line, "GAP-DECAY-LENGTH=%d", &i) == 1) { m->decay_gap_length = i; continue; }
    if (sscanf(line, "GAP-PENALTIES=%d %d %d", &i, &j, &k) == 3) {
      m->gap_set[0][0] = m->gap_set[1][0] = i;
      m->gap_set[0][1] = m->gap_set[1][1] = j;
      m->gap_set[0][2] = m->gap_set[1][2] = k;
      continue;
    }
    if (sscanf(line, "GAP-PENALTIES-X=%d %d %d", &i, &j, &k) == 3) {
      m->gap_set[1][0] = i; m->gap_set[1][1] = j; m->gap_set[1][2] = k;
      continue;
    }
    if (!have_symbols) {
      for (i = 0; line[i]; i++)
        if (!isspace((unsigned char)line[i]) && nsymb < ORA_MAXSYM) m->symbol[nsymb++] = line[i];
      have_symbols = 1;
      continue;
    }
    /* a score row: first char names the row symbol (last match wins the search) */
    have_symbols = 0;
    for (isymb = nsymb; isymb-- > 0;) {
      if (m->symbol[isymb] == line[0]) {
        have_symbols = 1;
        j = 1;
        for (i = 0; i < nsymb; i++) {
          if (sscanf(line + j, "%d%n", &m->score[isymb][i], &k) == 1) j += k;
          else { fclose(f); return -1; }
        }
        break;
      }
    }
    if (!have_symbols) { fclose(f); return -1; }
  }
  fclose(f);
  build_gap_tables(m);
  m->symbol[nsymb] = '\0';
  m->nsymbol = nsymb;
  return nsymb;
}

void ora_default_matrix(ora_matrix *m)
{
  static const char sym[] = "ARNDCQEGHILKMFPSTWYVBZX?agtcu]n";
  int i, j, n = (int)strlen(sym);
  memset(m, 0, sizeof *m);
  strcpy(m->symbol, sym);
  m->nsymbol = n;
  for (i = 0; i < n; i++)
    for (j = 0; j < n; j++) m->score[i][j] = (i == j) ? 0 : -10;
  m->gap_set[0][0] = m->gap_set[1][0] = 10;
  m->gap_set[0][1] = m->gap_set[1][1] = 5;
  m->gap_set[0][2] = m->gap_set[1][2] = 5;
  m->trunc_gap_length = 10;
  m->decay_gap_length = 5;
  build_gap_tables(m);
}

/* --------------------------------------------------------------- sequences */

void ora_index_sequence(const ora_matrix *m, const char *seq, int len, unsigned char *out)
{
  int i, j;
  for (i = 0; i < len; i++) {
    int c = tolower((unsigned char)seq[i]);
    int k;
    if (c == 0 || !strchr(m->symbol, c)) c = m->symbol[0]; /* limit_residues */
    k = m->nsymbol - 1;
    for (j = m->nsymbol; j-- > 0;)                          /* index_symbols: last match */
      if (m->symbol[j] == (char)c) { k = j; break; }
    out[i] = (unsigned char)k;
  }
}

static void po_alloc(ora_po *p, int n)
{
  int i;
  p->n = n;
  p->letter = (unsigned char *)calloc(n > 0 ? n : 1, 1);
  p->npred = (int *)calloc(n > 0 ? n : 1, sizeof(int));
  p->pred = (int *)calloc((size_t)(n > 0 ? n : 1) * ORA_MAXPRED, sizeof(int));
  p->src = (int *)malloc((size_t)(n > 0 ? n : 1) * ORA_MAXSRC * sizeof(int));
  p->ring_id = (int *)calloc(n > 0 ? n : 1, sizeof(int));
  p->align_ring = (int *)calloc(n > 0 ? n : 1, sizeof(int));
  for (i = 0; i < n * ORA_MAXSRC; i++) p->src[i] = -1;
  for (i = 0; i < n; i++) p->ring_id[i] = p->align_ring[i] = i;
}

void ora_po_free(ora_po *p)
{
  free(p->letter); free(p->npred); free(p->pred); free(p->src);
  free(p->ring_id); free(p->align_ring);
  memset(p, 0, sizeof *p);
}

void ora_po_linear(ora_po *p, const unsigned char *codes, int len)
{
  int i;
  po_alloc(p, len);
  p->nsrc = 1;
  p->src_len[0] = len;
  for (i = 0; i < len; i++) {
    p->letter[i] = codes[i];
    if (i > 0) { p->npred[i] = 1; p->pred[i * ORA_MAXPRED] = i - 1; }
    p->src[i * ORA_MAXSRC + 0] = i;
  }
}

/* ---------------------------------------------------------------- alignment */

#define NODE_INITIAL 1
#define NODE_FINAL 2

/* left list with the virtual -1 link: returns count, fills out[] */
static int left_list(const ora_po *p, int i, int type, int *out)
{
  int k, n = 0;
  if (p->npred[i] == 0) { out[0] = -1; return 1; }
  if (type & NODE_INITIAL) out[n++] = -1;
  for (k = 0; k < p->npred[i]; k++) out[n++] = p->pred[i * ORA_MAXPRED + k];
  return n;
}

static int *node_stats(const ora_po *p)
{
  int i, s;
  int *type = (int *)calloc(p->n > 0 ? p->n : 1, sizeof(int));
  for (i = 0; i < p->n; i++)
    for (s = 0; s < p->nsrc; s++) {
      int ipos = p->src[i * ORA_MAXSRC + s];
      if (ipos < 0) continue;
      if (ipos == 0) type[i] |= NODE_INITIAL;
      if (ipos == p->src_len[s] - 1) type[i] |= NODE_FINAL;
    }
  return type;
}

typedef struct { int score; short gx, gy; } cell_t;

int ora_align(const ora_po *x, const ora_po *y, const ora_matrix *m, int *x2y, int *y2x)
{
  int lx = x->n, ly = y->n, i, j, a, b;
  int M = m->max_gap_length;
  int *tx = node_stats(x), *ty = node_stats(y);
  int penx[64], peny[64], nxt[64];
  size_t W = (size_t)lx + 1;
  cell_t *S = (cell_t *)calloc(((size_t)ly + 1) * W, sizeof(cell_t));
  unsigned char *mvx = (unsigned char *)calloc((size_t)ly * lx + 1, 1);
  unsigned char *mvy = (unsigned char *)calloc((size_t)ly * lx + 1, 1);
  int best = MIN_SCORE, bx = -1, by = -1;
  int lxl[ORA_MAXPRED + 1], lyl[ORA_MAXPRED + 1], nlx, nly;
#define CELL(r, c) S[((size_t)(r) + 1) * W + (size_t)(c) + 1]

  for (i = 0; i <= M + 1; i++) { penx[i] = m->gap_penalty_x[i]; peny[i] = m->gap_penalty_y[i]; }
  for (i = 0; i < M + 1; i++) nxt[i] = (i < M) ? i + 1 : i;
  penx[M + 1] = penx[0]; peny[M + 1] = peny[0]; nxt[M + 1] = nxt[0]; /* global alignment */

  CELL(-1, -1).score = 0;
  CELL(-1, -1).gx = CELL(-1, -1).gy = (short)(M + 1);
  for (j = 0; j < lx; j++) {
    cell_t *c = &CELL(-1, j);
    c->score = MIN_SCORE;
    nlx = left_list(x, j, tx[j], lxl);
    for (a = 0; a < nlx; a++) {
      cell_t *p = &CELL(-1, lxl[a]);
      int t = p->score - penx[p->gx];
      if (t > c->score) { c->score = t; c->gx = (short)nxt[p->gx]; c->gy = (short)nxt[p->gx]; }
    }
  }
  for (i = 0; i < ly; i++) {
    cell_t *c = &CELL(i, -1);
    c->score = MIN_SCORE;
    nly = left_list(y, i, ty[i], lyl);
    for (b = 0; b < nly; b++) {
      cell_t *p = &CELL(lyl[b], -1);
      int t = p->score - peny[p->gy];
      if (t > c->score) { c->score = t; c->gx = (short)nxt[p->gy]; c->gy = (short)nxt[p->gy]; }
    }
  }

  for (i = 0; i < ly; i++) {
    nly = left_list(y, i, ty[i], lyl);
    for (j = 0; j < lx; j++) {
      int match = MIN_SCORE, mx = 0, my = 0;
      int insx = MIN_SCORE, ixx = 0, ixg = 0;
      int insy = MIN_SCORE, iyy = 0, iyg = 0;
      int end_ok = (tx[j] & NODE_FINAL) && (ty[i] & NODE_FINAL);
      cell_t *c = &CELL(i, j);
      nlx = left_list(x, j, tx[j], lxl);
      for (b = 0; b < nly; b++) {
        cell_t *p = &CELL(lyl[b], j);
        int t = p->score - peny[p->gy];
        if (t > insy) { insy = t; iyy = b + 1; iyg = p->gy; }
        for (a = 0; a < nlx; a++) {
          t = CELL(lyl[b], lxl[a]).score;
          if (t > match) { match = t; mx = a + 1; my = b + 1; }
        }
      }
      for (a = 0; a < nlx; a++) {
        cell_t *p = &CELL(i, lxl[a]);
        int t = p->score - penx[p->gx];
        if (t > insx) { insx = t; ixx = a + 1; ixg = p->gx; }
      }
      match += m->score[x->letter[j]][y->letter[i]];
      if (match > insy && match > insx) {
        c->score = match; c->gx = c->gy = 0;
        mvx[(size_t)i * lx + j] = (unsigned char)mx; mvy[(size_t)i * lx + j] = (unsigned char)my;
      } else if (insx > insy) {
        c->score = insx; c->gx = c->gy = (short)nxt[ixg];
        mvx[(size_t)i * lx + j] = (unsigned char)ixx; mvy[(size_t)i * lx + j] = 0;
      } else {
        c->score = insy; c->gx = c->gy = (short)nxt[iyg];
        mvx[(size_t)i * lx + j] = 0; mvy[(size_t)i * lx + j] = (unsigned char)iyy;
      }
      if (end_ok && c->score >= best)
        if (c->score > best || (j == bx && i < by) || j < bx) { best = c->score; bx = j; by = i; }
    }
  }

  for (j = 0; j < lx; j++) x2y[j] = -1;
  for (i = 0; i < ly; i++) y2x[i] = -1;
  while (bx >= 0 && by >= 0) {
    int xm = mvx[(size_t)by * lx + bx], ym = mvy[(size_t)by * lx + bx];
    if (xm > 0 && ym > 0) { x2y[bx] = by; y2x[by] = bx; }
    if (xm == 0 && ym == 0) { x2y[bx] = by; y2x[by] = bx; break; }
    if (xm > 0) { nlx = left_list(x, bx, tx[bx], lxl); }
    if (ym > 0) { nly = left_list(y, by, ty[by], lyl); }
    if (xm > 0) bx = lxl[xm - 1];
    if (ym > 0) by = lyl[ym - 1];
  }
#undef CELL
  free(S); free(mvx); free(mvy); free(tx); free(ty);
  return best;
}

/* -------------------------------------------------------------------- fusion */

static int reindex_fusion(const ora_po *x, const ora_po *y, const int *x2y, const int *y2x,
                          int *nx, int *ny)
{
  int lx = x->n, ly = y->n, ix, iy = 0, n = 0, ir, end_of_ring = -1;
  char *fuse = (char *)calloc(ly > 0 ? ly : 1, 1);
  for (ir = 0; ir < ly; ir++)
    if (y2x[ir] >= 0 && x->letter[y2x[ir]] == y->letter[ir]) fuse[ir] = 1;
  for (ix = 0; ix < lx; ix++) {
    for (ir = ix; ir < lx && x->ring_id[ir] == x->ring_id[ix]; ir++)
      if (x2y[ir] >= 0) {
        while (iy < x2y[ir]) ny[iy++] = n++;
        break;
      }
    if (x2y[ix] >= 0 && iy < ly) {
      for (ir = y->align_ring[iy]; ir != iy; ir = y->align_ring[ir])
        if (ir > end_of_ring) end_of_ring = ir;
      if (fuse[iy]) ny[iy++] = n;
      else ny[iy++] = n++;
    }
    nx[ix] = n++;
    while (iy <= end_of_ring) ny[iy++] = n++;
  }
  while (iy < ly) ny[iy++] = n++;
  free(fuse);
  return n;
}

static void add_link(ora_po *p, int node, int to)
{
  int k;
  for (k = 0; k < p->npred[node]; k++)
    if (p->pred[node * ORA_MAXPRED + k] == to) return;
  if (p->npred[node] >= ORA_MAXPRED) { fprintf(stderr, "oracle: ORA_MAXPRED exceeded\n"); abort(); }
  p->pred[node * ORA_MAXPRED + p->npred[node]++] = to;
}

static void crosslink(ora_po *p, int a, int b)
{
  int r, t;
  if (p->ring_id[a] == p->ring_id[b]) return;
  if (p->ring_id[a] < p->ring_id[b]) {
    r = b;
    do p->ring_id[r] = p->ring_id[a]; while ((r = p->align_ring[r]) != b);
  } else {
    r = a;
    do p->ring_id[r] = p->ring_id[b]; while ((r = p->align_ring[r]) != a);
  }
  t = p->align_ring[a]; p->align_ring[a] = p->align_ring[b]; p->align_ring[b] = t;
}

void ora_fuse(ora_po *x, const ora_po *y, const int *x2y, const int *y2x)
{
  int lx = x->n, ly = y->n, i, k, s, new_len;
  int *nx = (int *)calloc(lx > 0 ? lx : 1, sizeof(int));
  int *ny = (int *)calloc(ly > 0 ? ly : 1, sizeof(int));
  ora_po z;
  char *is_x;

  new_len = reindex_fusion(x, y, x2y, y2x, nx, ny);
  po_alloc(&z, new_len);
  is_x = (char *)calloc(new_len > 0 ? new_len : 1, 1);
  z.nsrc = x->nsrc + y->nsrc;
  for (s = 0; s < x->nsrc; s++) z.src_len[s] = x->src_len[s];
  for (s = 0; s < y->nsrc; s++) z.src_len[x->nsrc + s] = y->src_len[s];

  /* x letters move to nx[] with every index translated (translate_lpo) */
  for (i = 0; i < lx; i++) {
    int d = nx[i];
    is_x[d] = 1;
    z.letter[d] = x->letter[i];
    z.npred[d] = x->npred[i];
    for (k = 0; k < x->npred[i]; k++) z.pred[d * ORA_MAXPRED + k] = nx[x->pred[i * ORA_MAXPRED + k]];
    for (s = 0; s < x->nsrc; s++) z.src[d * ORA_MAXSRC + s] = x->src[i * ORA_MAXSRC + s];
    z.ring_id[d] = nx[x->ring_id[i]];
    z.align_ring[d] = nx[x->align_ring[i]];
  }
  /* y letters: copy_lpo_letter, last to first */
  for (i = ly; i-- > 0;) {
    int d = ny[i];
    z.letter[d] = y->letter[i];
    for (s = 0; s < y->nsrc; s++)
      if (y->src[i * ORA_MAXSRC + s] >= 0) z.src[d * ORA_MAXSRC + x->nsrc + s] = y->src[i * ORA_MAXSRC + s];
    for (k = 0; k < y->npred[i]; k++) add_link(&z, d, ny[y->pred[i * ORA_MAXPRED + k]]);
  }
  for (i = ly; i-- > 0;) { /* copy_old_ring_to_new for y */
    int ipos, nextp;
    for (ipos = i; (nextp = y->align_ring[ipos]) != i; ipos = nextp) crosslink(&z, ny[ipos], ny[nextp]);
  }
  for (i = lx; i-- > 0;)
    if (x2y[i] >= 0) crosslink(&z, nx[i], ny[x2y[i]]);

  free(nx); free(ny); free(is_x);
  ora_po_free(x);
  *x = z;
}

/* ---------------------------------------------------------------------- emit */

int ora_emit(const ora_po *p, const ora_matrix *m, char **rows_out)
{
  int i, s, nring = 0, cur = 0, ir = 0;
  char *rows;
  for (i = 0; i < p->n; i++)
    if (p->ring_id[i] != cur) { cur = p->ring_id[i]; nring++; }
  nring++;
  rows = (char *)malloc((size_t)p->nsrc * nring + 1);
  memset(rows, '.', (size_t)p->nsrc * nring);
  cur = 0;
  for (i = 0; i < p->n; i++) {
    if (p->ring_id[i] != cur) { cur = p->ring_id[i]; ir++; }
    for (s = 0; s < p->nsrc; s++)
      if (p->src[i * ORA_MAXSRC + s] >= 0)
        rows[(size_t)s * nring + ir] = (p->letter[i] < m->nsymbol) ? m->symbol[p->letter[i]] : (char)p->letter[i];
  }
  *rows_out = rows;
  return nring;
}

/* -------------------------------------------------------------------- window */

int ora_window(const ora_matrix *m, const char *ref, int lr, const char *cor, int lc,
               const char *unc, int lu, ora_result *res)
{
  unsigned char *cr = (unsigned char *)malloc(lr + 1), *cc = (unsigned char *)malloc(lc + 1),
                *cu = (unsigned char *)malloc(lu + 1);
  ora_po P, C, U;
  memset(res, 0, sizeof *res);
  if (lr <= 0 || lc <= 0 || lu <= 0) { free(cr); free(cc); free(cu); return -1; }
  ora_index_sequence(m, ref, lr, cr);
  ora_index_sequence(m, cor, lc, cc);
  ora_index_sequence(m, unc, lu, cu);
  ora_po_linear(&P, cr, lr);
  ora_po_linear(&C, cc, lc);
  ora_po_linear(&U, cu, lu);
  res->x2y1 = (int *)malloc(sizeof(int) * lr);
  res->y2x1 = (int *)malloc(sizeof(int) * lc);
  res->score1 = ora_align(&P, &C, m, res->x2y1, res->y2x1);
  ora_fuse(&P, &C, res->x2y1, res->y2x1);
  res->len_p1 = P.n;
  res->x2y2 = (int *)malloc(sizeof(int) * P.n);
  res->y2x2 = (int *)malloc(sizeof(int) * lu);
  res->score2 = ora_align(&P, &U, m, res->x2y2, res->y2x2);
  ora_fuse(&P, &U, res->x2y2, res->y2x2);
  res->len_p2 = P.n;
  res->cells = (long long)lr * lc + (long long)res->len_p1 * lu;
  res->nring = ora_emit(&P, m, &res->rows);
  ora_po_free(&P); ora_po_free(&C); ora_po_free(&U);
  free(cr); free(cc); free(cu);
  return 0;
}

void ora_free_result(ora_result *res)
{
  free(res->rows); free(res->x2y1); free(res->y2x1); free(res->x2y2); free(res->y2x2);
  memset(res, 0, sizeof *res);
}

int ora_batch(const ora_matrix *m, int n, const char *ref, const long long *ref_off,
              const char *cor, const long long *cor_off, const char *unc,
              const long long *unc_off, char *rows_out, long long *row_off, int *nring,
              int *score1, int *score2, long long *cells, int nthreads)
{
  int w;
  long long acc = 0;
  /* every window's MSA has at most lr+lc+lu columns: reserve 3x that per window */
  for (w = 0; w < n; w++) {
    row_off[w] = acc;
    acc += 3 * ((ref_off[w + 1] - ref_off[w]) + (cor_off[w + 1] - cor_off[w]) + (unc_off[w + 1] - unc_off[w]));
  }
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (w = 0; w < n; w++) {
    ora_result r;
    int lr = (int)(ref_off[w + 1] - ref_off[w]), lc = (int)(cor_off[w + 1] - cor_off[w]),
        lu = (int)(unc_off[w + 1] - unc_off[w]);
    if (ora_window(m, ref + ref_off[w], lr, cor + cor_off[w], lc, unc + unc_off[w], lu, &r) != 0) {
      nring[w] = 0; score1[w] = score2[w] = 0; if (cells) cells[w] = 0;
      continue;
    }
    memcpy(rows_out + row_off[w], r.rows, (size_t)3 * r.nring);
    nring[w] = r.nring; score1[w] = r.score1; score2[w] = r.score2;
    if (cells) cells[w] = r.cells;
    ora_free_result(&r);
  }
  return 0;
}

/* --------------------------------------------------------------- file driver */

typedef struct { char *name, *title, *seq; int len; } fa_rec;
typedef struct { fa_rec *r; int n, cap; } fa_list;

#define FA_LINE 32768 /* SEQ_LENGTH_MAX, seq_util.h:22: longer lines arrive in chunks */
#define FA_NAME 4096
#define NAME_MAX_KEEP 512 /* SEQUENCE_NAME_MAX: names are cut to 511 chars */

static void fa_push(fa_list *L, const char *name, const char *title, char *buf)
{
  int i, j = 0;
  fa_rec *r;
  if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 64; L->r = (fa_rec *)realloc(L->r, L->cap * sizeof(fa_rec)); }
  r = &L->r[L->n++];
  for (i = 0; buf[i]; i++)
    if (!isspace((unsigned char)buf[i])) buf[j++] = buf[i];
  buf[j] = '\0';
  r->seq = strdup(buf);
  r->len = j;
  r->name = strdup(name);
  if ((int)strlen(r->name) > NAME_MAX_KEEP - 1) r->name[NAME_MAX_KEEP - 1] = '\0';
  r->title = strdup(title);
}

static int fa_read(const char *path, fa_list *L)
{
  static char line[FA_LINE];
  char *name = (char *)calloc(FA_LINE + 8, 1), *title = (char *)calloc(FA_LINE + 8, 1);
  char *buf = NULL; size_t blen = 0, bcap = 0;
  char *p; int c;
  FILE *f = fopen(path, "r");
  memset(L, 0, sizeof *L);
  if (!f) { free(name); free(title); return -1; }
  while (fgets(line, sizeof(line) - 1, f)) {
    if ((p = strrchr(line, '\n'))) *p = '\0';
    switch (line[0]) {
    case '#': break;
    case '>':
      if (name[0] && buf && buf[0]) fa_push(L, name, title, buf);
      name[0] = '\0';
      if (sscanf(line + 1, "%s %[^\n]", name, title) < 2) strcpy(title, "untitled");
      if (buf) buf[0] = '\0';
      blen = 0;
      break;
    case '*': break;
    default:
      if (name[0]) {
        size_t l = strlen(line);
        if (blen + l + 1 > bcap) { bcap = 2 * (blen + l + 1) + 4096; buf = (char *)realloc(buf, bcap); }
        memcpy(buf + blen, line, l + 1);
        blen += l;
      }
    }
    c = getc(f);
    if (c == EOF) break;
    ungetc(c, f);
    if (c == '#' && L->n > 0) break;
  }
  if (name[0] && buf && buf[0]) fa_push(L, name, title, buf);
  fclose(f);
  free(buf); free(name); free(title);
  return L->n;
}

static void fa_free(fa_list *L)
{
  int i;
  for (i = 0; i < L->n; i++) { free(L->r[i].name); free(L->r[i].title); free(L->r[i].seq); }
  free(L->r);
}

int ora_poa_files(const char *matrix, const char *ref_fa, const char *cor_fa,
                  const char *unc_fa, const char *pir_out, int print_perm)
{
  ora_matrix m;
  fa_list R, C, U;
  FILE *out;
  int i, s, n;
  if (ora_read_matrix(matrix, &m) <= 0) return 1;
  if (fa_read(cor_fa, &C) < 0) return 1;
  if (fa_read(unc_fa, &U) < 0) { fa_free(&C); return 1; }
  if (fa_read(ref_fa, &R) < 0) { fa_free(&C); fa_free(&U); return 1; }
  n = R.n;
  if (n == 0) { fa_free(&R); fa_free(&C); fa_free(&U); return 1; }
  if (C.n < n) n = C.n; /* the reference reads past the shorter arrays (undefined); stop cleanly */
  if (U.n < n) n = U.n;
  out = fopen(pir_out, "w");
  if (!out) { fa_free(&R); fa_free(&C); fa_free(&U); return 1; }
  for (i = 0; i < n; i++) {
    ora_result r;
    fa_rec *recs[3];
    recs[0] = &R.r[i]; recs[1] = &C.r[i]; recs[2] = &U.r[i];
    ora_window(&m, R.r[i].seq, R.r[i].len, C.r[i].seq, C.r[i].len, U.r[i].seq, U.r[i].len, &r);
    if (print_perm) printf("0 1 2 \n");
    for (s = 0; s < 3; s++) {
      fprintf(out, ">%s %s\n", recs[s]->name, recs[s]->title);
      fwrite(r.rows + (size_t)s * r.nring, 1, r.nring, out);
      fputc('\n', out);
    }
    ora_free_result(&r);
  }
  fclose(out);
  n = (R.n == C.n && R.n == U.n) ? 0 : 1;
  fa_free(&R); fa_free(&C); fa_free(&U);
  return n;
}

/* same text format as oracle/ref_harness.c (the dump of the compiled reference) */
static void dump_map(FILE *f, const char *tag, int n, const int *m)
{
  int i;
  fprintf(f, "%s %d", tag, n);
  for (i = 0; i < n; i++) fprintf(f, " %d", m[i]);
  fputc('\n', f);
}

int ora_dump_files(const char *matrix, const char *ref_fa, const char *cor_fa,
                   const char *unc_fa, const char *dump_out)
{
  ora_matrix m;
  fa_list R, C, U;
  FILE *out;
  int i, n;
  if (ora_read_matrix(matrix, &m) <= 0) return 1;
  if (fa_read(cor_fa, &C) < 0) return 1;
  if (fa_read(unc_fa, &U) < 0) { fa_free(&C); return 1; }
  if (fa_read(ref_fa, &R) < 0) { fa_free(&C); fa_free(&U); return 1; }
  n = R.n;
  if (C.n < n) n = C.n;
  if (U.n < n) n = U.n;
  out = fopen(dump_out, "w");
  if (!out) return 1;
  for (i = 0; i < n; i++) {
    ora_result r;
    ora_window(&m, R.r[i].seq, R.r[i].len, C.r[i].seq, C.r[i].len, U.r[i].seq, U.r[i].len, &r);
    fprintf(out, "W %d %d %d %d\n", i, R.r[i].len, C.r[i].len, U.r[i].len);
    fprintf(out, "S1 %d\n", r.score1);
    dump_map(out, "X1", R.r[i].len, r.x2y1);
    dump_map(out, "Y1", C.r[i].len, r.y2x1);
    fprintf(out, "S2 %d %d\n", r.score2, r.len_p1);
    dump_map(out, "X2", r.len_p1, r.x2y2);
    dump_map(out, "Y2", U.r[i].len, r.y2x2);
    fprintf(out, "L %d\n", r.len_p2);
    ora_free_result(&r);
  }
  fclose(out);
  fa_free(&R); fa_free(&C); fa_free(&U);
  return 0;
}

#ifdef ORA_MAIN
/* oracle CLI with the reference's five flags (main.c:85-113) */
int main(int argc, char **argv)
{
  const char *pir = "default_output_msa.fasta", *cor = NULL, *unc = NULL, *ref = NULL, *mat = "./blosum80.mat";
  int i;
  if (argc < 2) { fprintf(stderr, "usage: %s -pir OUT -corrected_reads_fasta F -uncorrected_reads_fasta F -reference_reads_fasta F -pathMatrix M\n", argv[0]); exit(-1); }
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-pir")) { pir = argv[++i]; continue; }
    if (!strcmp(argv[i], "-corrected_reads_fasta")) { cor = argv[++i]; continue; }
    if (!strcmp(argv[i], "-uncorrected_reads_fasta")) { unc = argv[++i]; continue; }
    if (!strcmp(argv[i], "-reference_reads_fasta")) { ref = argv[++i]; continue; }
    if (!strcmp(argv[i], "-pathMatrix")) { mat = argv[++i]; continue; }
  }
  if (!(cor && unc && ref)) return 0;
  return ora_poa_files(mat, ref, cor, unc, pir, 1);
}
#endif

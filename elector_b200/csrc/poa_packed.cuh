// poa_packed.cuh -- the 16-bit packed variant of the DP kernels (the one the bulk of the windows runs).
//
// poa_kernel.cuh computes one DP cell per ~15 INT32 instructions and its bulk launches sit at
// 60-77 % ALU-pipe utilisation (profiles/r1c_*): the integer issue rate is the limit.  sm_100a has
// packed halfword integer instructions (VIMNMX.S16x2 with predicate outputs, VIADDMNMX.S16x2,
// VIMNMX.U16x2); this file computes TWO cells per instruction with them.
//
// Two cells of one window can only be updated together if they do not depend on each other.  A
// band of 2R rows (R = 6, 7 or 8) is split into a low half (rows r0 .. r0+R-1) and a high half
// (rows r0+R .. r0+2R-1) that runs ONE COLUMN BEHIND: iteration j updates (r0+k, j) in the low
// 16 bits and (r0+R+k, j-1) in the high 16 bits of the same registers.  The high half's first row
// takes its "up" and "diagonal" inputs from the low half's last row of the previous iterations.
// Iteration 0 makes the high half reproduce the virtual column -1 by itself (its registers start
// at 0 = minus infinity, so every cell takes the Y-gap move from the cell above), and the last
// iteration (j = len_x) only completes the high half; no masking is needed anywhere.
//
// Arithmetic.  Halves hold S + kBiasP (always in [1, 32767]), so plain 32-bit add / subtract of
// packed operands never borrows across the halves where it is used.  Per pair of cells:
//   eq  = x2 ^ y2[k]                       letters (code << 4) differ  <=>  half >= 16
//   t   = min.u16x2(eq, |mismatch|)        substitution penalty
//   M   = diag - t
//   gap = max.s16x2(up, pG)  -> predicates (pG > up): the Y-gap wins ties     (align_lpo_po2.c:392)
//   s   = max.s16x2(gap, M)  -> predicates (M > gap): a match must beat both  (:384)
//   g   = max.s16x2(M - open, gap - ext)   = s - pen(move)      [see packed_ok()]
//   4 predicated ORs set the two move bits of the two cells
// = 11 instructions for 2 cells.  Exact for matrices in the class packed_ok() describes (the shipped
// blosum80.mat is); every other matrix, and windows whose scores could leave 16 bits, run the
// INT32 kernels of poa_kernel.cuh.
#pragma once
#include "poa_kernel.cuh"

namespace elector {

constexpr int kBiasP = 16384;       // halves hold S + kBiasP
constexpr int kNegP = kBiasP - 8000; // minus infinity of the diagonal band (biased): see BandW in poa_kernel.cuh
constexpr int kPackedSpan = 16000;  // a segment runs packed when maxabs * (len_x + len_y + 4) <= kPackedSpan

// rows per half-band of a group whose longest row sequence has ly letters: ceil(ly / 16) bands of 2R rows
EL_HD int packed_rows(int ly) {
  const int nb = (ly + 15) >> 4;
  const int need = (ly + 2 * nb - 1) / (2 * nb);
  return need <= 6 ? 6 : need <= 7 ? 7 : 8;
}

// ---- packed halfword primitives (single instructions on sm_100a; plain C on the host for tests/emul) ----
EL_HD uint32_t pk_minu(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vminu2(a, b);
#else
  const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
  return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
// per-half signed max(a, b); ORs bit_lo / bit_hi into mv where b beats a (b > a) in the low / high half.
// The PTX below is the pattern ptxas turns into ONE VIMNMX.S16x2 with two predicate outputs (the same
// pattern as CUDA's __vibmax_s16x2) followed by two predicated LOP3.
EL_HD uint32_t pk_maxs_flag(uint32_t a, uint32_t b, uint32_t &mv, uint32_t bit_lo, uint32_t bit_hi) {
#ifdef __CUDA_ARCH__
  uint32_t val;
  asm("{.reg .pred pu, pv;\n\t"
      ".reg .s16 rs0, rs1, rs2, rs3;\n\t"
      "max.s16x2 %0, %2, %3;\n\t"
      "mov.b32 {rs0, rs1}, %0;\n\t"
      "mov.b32 {rs2, rs3}, %2;\n\t"
      "setp.eq.s16 pv, rs0, rs2;\n\t"
      "setp.eq.s16 pu, rs1, rs3;\n\t"
      "@!pv or.b32 %1, %1, %4;\n\t"
      "@!pu or.b32 %1, %1, %5;}\n\t"
      : "=&r"(val), "+r"(mv) : "r"(a), "r"(b), "r"(bit_lo), "r"(bit_hi));   // early clobber: a is read after val is written
  return val;
#else
  const int16_t al = (int16_t)(a & 0xffffu), bl = (int16_t)(b & 0xffffu), ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
  if (bl > al) mv |= bit_lo;
  if (bh > ah) mv |= bit_hi;
  return (uint32_t)(uint16_t)(al >= bl ? al : bl) | ((uint32_t)(uint16_t)(ah >= bh ? ah : bh) << 16);
#endif
}
// per-half signed max(a + b, c)
EL_HD uint32_t pk_addmaxs(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __viaddmax_s16x2(a, b, c);
#else
  const int16_t sl = (int16_t)((a + b) & 0xffffu), sh = (int16_t)(((a >> 16) + (b >> 16)) & 0xffffu);
  const int16_t cl = (int16_t)(c & 0xffffu), ch = (int16_t)(c >> 16);
  return (uint32_t)(uint16_t)(sl > cl ? sl : cl) | ((uint32_t)(uint16_t)(sh > ch ? sh : ch) << 16);
#endif
}
EL_HD uint32_t pk2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
EL_HD uint32_t pk_lo_hi(uint32_t lo_src, uint32_t hi_src) {   // low half of lo_src, high half of hi_src
  return (lo_src & 0xffffu) | (hi_src & 0xffff0000u);
}

struct PackedConsts {
  uint32_t mis2, nopen2, ext2;   // |mismatch|, -open, ext in both halves
  EL_HD void set(const Scoring &sc) {
    mis2 = pk2(-sc.mismatch, -sc.mismatch);
    nopen2 = pk2(-sc.open, -sc.open);
    ext2 = pk2(sc.ext, sc.ext);
  }
};

// The packed update of one iteration: the low halves of S/G move from column j-1 to column j, the
// high halves from the column before to the column of the previous iteration.  Returns the move
// bits: cell (half h, row k) at bits 16h + 2k + 1 (match) and 16h + 2k (X-gap when not a match).
template <int R>
EL_HD uint32_t update_packed(const PackedConsts &pc, uint32_t (&S)[R], uint32_t (&G)[R], const uint32_t (&y2)[R], uint32_t x2,
                             uint32_t diag, uint32_t up) {
  uint32_t mv = 0;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t pS = S[k], pG = G[k];
    const uint32_t t = pk_minu(x2 ^ y2[k], pc.mis2);
    const uint32_t M = diag - t;
    const uint32_t gap = pk_maxs_flag(up, pG, mv, 1u << (2 * k), 1u << (16 + 2 * k));   // X-gap only when it beats the Y-gap
    const uint32_t s = pk_maxs_flag(gap, M, mv, 2u << (2 * k), 2u << (16 + 2 * k));     // match only when it beats both
    const uint32_t g = pk_addmaxs(M, pc.nopen2, gap - pc.ext2);
    S[k] = s; G[k] = g;
    diag = pS; up = g;
  }
  return mv;
}

template <int R>
EL_HD int pick_half(const uint32_t (&S)[R], int k, bool hi) {
  uint32_t s = S[0];
#pragma unroll
  for (int r = 1; r < R; ++r) if (k == r) s = S[r];
  return (int)(hi ? s >> 16 : s & 0xffffu) - kBiasP;
}

// =============================== phase 1, packed ===============================================
// node record: node j lives in record j + 1 (record 0 = the virtual column -1, which the high half
// produces in iteration 0 like any other column)
enum : uint32_t { P1_BSG = 0, P1_X2Y = 1, P1_MOVES = 2 };   // boundary S | G << 16 below the band, x2y, one moves word per band

struct Layout1P {
  uint32_t o_ref, o_cor, o_nodes, rec_words, R, total;
};
EL_HD void make_layout1p(Layout1P &L, int LR, int LC) {
  uint32_t o = 0;
  L.o_ref = o; o += cdiv_u(LR, 4) + 3;                         // the last iteration reads one letter past the end, the look-ahead one word more
  L.o_cor = o; o += cdiv_u(LC, 4) + 5;                         // rows up to 16 * ceil(LC / 16) - 1 are read
  L.R = (uint32_t)packed_rows(LC);
  L.rec_words = P1_MOVES + cdiv_u(LC, 16);
  L.o_nodes = o; o += ((uint32_t)LR + 4) * L.rec_words;        // record 0, LR nodes, three records of look-ahead
  L.total = o;
}

struct Phase1P {
  typedef Layout1P Layout;
  static constexpr bool kGenericSub = false;
  static constexpr bool kBanded = true;
  LaneScratch scr;
  Scoring sc;
  const Layout1P *Lp;
  BandW bw = {0, 0, 0, false};   // set per group by the kernel
  static EL_HD void make_layout(Layout1P &L, int LR, int LC) { make_layout1p(L, LR, LC); }

  EL_HD uint32_t *rec(int j) const { return scr.at(Lp->o_nodes + (uint32_t)(j + 1) * Lp->rec_words); }
  // the band-restricted DP of a window is exact when it ends above the best score a path leaving the band can have
  // (the bound itself must lie above the band's minus infinity: out-of-band cells then never beat an in-band path that passes the test)
  EL_HD bool band_exact(int score, int lx, int ly) const {
    if (!bw.on) return true;
    const int bound = band_bound(sc, bw.w, ly - lx);
    return score > bound && bound > kNegP - kBiasP;
  }

  // one band of 2R rows of DP1 (lin(ref) columns x lin(cor) rows); returns the score of the last cell
  // when this is the band that holds row ly - 1
  template <int R>
  EL_HDN int band(int lr, int ly, int b, bool last, const BandW &bw) const {
    const int r0 = b * 2 * R;
    PackedConsts pc;
    pc.set(sc);
    // columns this band sweeps: all of them, or those whose offsets to the band's rows lie in [bw.omin, bw.omax]
    int jlo = 0, jend = lr;                                    // iterations jlo .. jend (the last one only completes the high half)
    if (bw.on) {
      jlo = r0 - bw.omax > 0 ? r0 - bw.omax : 0;
      const int jhi = r0 + 2 * R - 1 - bw.omin;
      jend = jhi + 1 < lr ? jhi + 1 : lr;
      if (jlo > jend) jlo = jend;
    }
    uint32_t y2[R], S[R], G[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      y2[k] = ((uint32_t)scr.code_at(Lp->o_cor, r0 + k) | ((uint32_t)scr.code_at(Lp->o_cor, r0 + R + k) << 16)) << 4;
      const int v = kBiasP + sc.virt_S(r0 + k);                // virtual column -1 (align_lpo_po2.c:290-302); high half: -inf
      S[k] = (uint32_t)v;
      G[k] = (uint32_t)(v - sc.ext);
      if (jlo > 0) S[k] = G[k] = pk2(kNegP, kNegP);            // the column before the band's first one is outside the band
    }
    const int rr = ly - 1 - r0;                                // row of the last cell inside this band (when last)
    uint32_t *p = rec(jlo - 1);
    const uint32_t step = Lp->rec_words * 32;
    // boundary row r0 - 1 at nodes j-1 / j: S in the low half, G in the high half
    // (loads run TWO iterations ahead of their use: the scratch of a launch exceeds the L2 and the long-scoreboard stall
    // was the top stall reason of this loop, profiles/r1f)
    uint32_t bsg, bsg_n, bsg_n2 = 0;
    if (b == 0) {                                              // row -1 (:272-286): corner, then -(open + ext * j)
      bsg = pk2(kBiasP, kBiasP - sc.open);
      bsg_n = pk2(kBiasP - sc.open, kBiasP - sc.open - sc.ext);
    } else {
      bsg = p[P1_BSG * 32];
      bsg_n = p[step + P1_BSG * 32];
      bsg_n2 = p[2 * step + P1_BSG * 32];
    }
    uint32_t xw = 0, xw_next = scr.w(Lp->o_ref + (jlo >> 2)), x2 = 0, d7 = 0, mlo = 0;
    if (jlo > 0) { x2 = (uint32_t)scr.code_at(Lp->o_ref, jlo - 1) << 4; d7 = (uint32_t)kNegP; }
    int best = 0;
    for (int j = jlo; j <= jend; ++j, p += step) {             // p = record of node j - 1
      if ((j & 3) == 0 || j == jlo) { xw = xw_next >> (8 * (j & 3)); xw_next = scr.w(Lp->o_ref + (j >> 2) + 1); }   // the next letters, one group ahead
      x2 = (x2 << 16) | ((xw & 0xffu) << 4);
      xw >>= 8;
      const uint32_t bsg_p = bsg;
      bsg = bsg_n;
      if (b == 0) bsg_n = bsg - pc.ext2;
      else { bsg_n = bsg_n2; bsg_n2 = p[3 * step + P1_BSG * 32]; }   // node j + 2
      const uint32_t diag0 = pk_lo_hi(bsg_p, d7 << 16);        // S(r0-1, j-1) | S(r0+R-1, j-2)
      const uint32_t up0 = (bsg >> 16) | (G[R - 1] << 16);     // G(r0-1, j)   | G(r0+R-1, j-1)
      d7 = S[R - 1];
      const uint32_t mv = update_packed<R>(pc, S, G, y2, x2, diag0, up0);
      // node j - 1 is now complete in the high half
      if (!last) p[P1_BSG * 32] = (S[R - 1] >> 16) | (G[R - 1] & 0xffff0000u);
      p[(P1_MOVES + b) * 32] = pk_lo_hi(mlo, mv);
      mlo = mv;
      if (last && j == lr - 1 && rr < R) best = pick_half<R>(S, rr, false);
    }
    if (last && rr >= R) best = pick_half<R>(S, rr - R, true);
    if (bw.on && !last) {                                      // the next band reads boundary rows 2R + 3 nodes further: outside this band
      const int stop = jend + 2 * R + 3 < lr + 2 ? jend + 2 * R + 3 : lr + 2;
      for (int j = jend; j <= stop; ++j, p += step) p[P1_BSG * 32] = pk2(kNegP, kNegP);   // p = record of node j
    }
    return best;
  }

  template <int R>
  EL_HDN int dp(int lr, int ly, const BandW &bw) const {
    const int nb = (ly + 2 * R - 1) / (2 * R);
    for (int b = 0; b < nb - 1; ++b) band<R>(lr, ly, b, false, bw);
    return band<R>(lr, ly, nb - 1, true, bw);
  }

  // traceback (align_lpo_po2.c:108-168): fills the x2y field of every record
  template <int R>
  EL_HDN void traceback(int lr, int ly) const {
    const ptrdiff_t step = (ptrdiff_t)Lp->rec_words * 32;
    {
      uint32_t *p = rec(0) + P1_X2Y * 32;
      for (int j = 0; j < lr; ++j, p += step) *p = 0xffffffffu;
    }
    int j = lr - 1, r = ly - 1;
    while (j >= 0 && r >= 0) {
      const int b = r / (2 * R);
      int rr = r - b * 2 * R;
      uint32_t *p = rec(j);
      const uint32_t *pm = p + (P1_MOVES + b) * 32;
      uint32_t w0 = pm[0], w1 = j >= 1 ? pm[-step] : 0, w2 = j >= 2 ? pm[-2 * step] : 0, w3 = j >= 3 ? pm[-3 * step] : 0;
      for (;;) {
        const uint32_t kind = (w0 >> (rr < R ? 2 * rr : 16 + 2 * (rr - R))) & 3u;   // bit 1 match, bit 0 X-gap
        if (kind & 2u) p[P1_X2Y * 32] = (uint32_t)r;
        if (kind != 1u) { --r; --rr; }
        if (kind) {
          --j; p -= step; pm -= step;
          w0 = w1; w1 = w2; w2 = w3;
          w3 = j >= 3 ? pm[-3 * step] : 0;
        }
        if (j < 0 || r < 0 || rr < 0) break;
      }
    }
  }

  template <int R>
  EL_HDN int run_r(int lr, int lc, uint16_t *p1_out, int &s1, int &spcode, bool &exact) const {
    s1 = dp<R>(lr, lc, bw);
    exact = band_exact(s1, lr, lc);
    if (!exact) return 0;                                      // to be run again without a band
    traceback<R>(lr, lc);
    return fuse1(scr, Lp->o_ref, Lp->o_cor, rec(0) + P1_X2Y * 32, (ptrdiff_t)Lp->rec_words * 32, lr, lc, p1_out, spcode);
  }

  // exact = false (band on only): nothing was written to p1_out, the window has to be run again without a band
  EL_HDN int run_window(const uint8_t *ref, int lr, const uint8_t *cor, int lc, uint16_t *p1_out, int &s1, int &spcode, bool &exact) const {
    scr.pack_codes(sc.tab, ref, lr, Lp->o_ref);
    scr.pack_codes(sc.tab, cor, lc, Lp->o_cor);
    // the last band reads up to 15 letters past the end of cor, the last iteration one past the end of ref:
    // defined values (which ones does not matter, they only feed cells outside the window)
    for (uint32_t k = 0; k < 5; ++k) scr.w(Lp->o_cor + cdiv_u((uint32_t)lc, 4) + k) = 0;
    scr.w(Lp->o_ref + cdiv_u((uint32_t)lr, 4)) = 0;
    if (Lp->R == 6) return run_r<6>(lr, lc, p1_out, s1, spcode, exact);   // R is uniform over the warp
    if (Lp->R == 7) return run_r<7>(lr, lc, p1_out, s1, spcode, exact);
    return run_r<8>(lr, lc, p1_out, s1, spcode, exact);
  }
};

// =============================== phase 2, packed ===============================================
// DP2 (P1 columns x lin(unc) rows) with the same skewed half-bands: iteration j updates node j in
// the low halves and node j-1 in the high halves.  The two frontier sets of poa_kernel.cuh become
// two sets of packed registers / shared-memory words whose halves are one node apart; what an
// uncommon node does to the sets (swap, copy, merge with the virtual column, align_lpo_po2.c:334-371)
// depends only on the node's shape and on which frontiers the sets hold, never on the scores, so
// the high half simply REPLAYS, one iteration later, what the low half did (arrange_half).
//
// node record: node j lives in record j + 1 (record 0 takes the stores of iteration 0's idle high half)
enum : uint32_t { Q2_NODE = 0, Q2_BSG = 1, Q2_PRED = 2, Q2_X2Y = 3, Q2_MOVES = 4 };   // BSG: boundary S | G << 16 below the band

constexpr int kSetWordsP = (2 * 8 + 1) * 32;   // one packed frontier set of a warp: S[R], G[R], h (low half only), lane-interleaved

struct Layout2P {
  uint32_t o_unc, o_nodes, rec_words, o_ord, ord_bands, o_nt, o_rows, row_words, R, total;
};
EL_HD void make_layout2p(Layout2P &L, int N1, int LU) {
  const uint32_t nb = cdiv_u(LU, 16);
  uint32_t o = 0;
  L.o_unc = o; o += cdiv_u(LU, 4) + 5;                          // rows up to 16 * ceil(LU / 16) - 1 are read
  L.R = (uint32_t)packed_rows(LU);
  L.rec_words = Q2_MOVES + nb;
  L.o_nodes = o; o += ((uint32_t)N1 + 2) * L.rec_words;         // record 0, N1 nodes, one record of look-ahead
  L.ord_bands = nb;
  L.o_ord = o; o += ((uint32_t)N1 / 2 + 2) * nb * 2;            // two words per (combined node, band): winning predecessor ordinals
  L.o_nt = o; o += cdiv_u(N1, 32) + 1;
  L.row_words = cdiv_u(N1 + LU, 4);
  L.o_rows = o; o += 3 * L.row_words;
  L.total = o;
}

// One half (low or high 16 bits) of the two frontier sets at an uncommon node; the counterpart of
// arrange_sets() in poa_kernel.cuh.  wa / wb = the lane's words of sets A / B (element k at [k * 32]):
// S[0..R-1], G[0..R-1] and, in the low half only, h = S of the row above the band.  row0 = DP row of
// element 0, ordrow0 = its row inside the band.  On return set A holds the node's source column and
// set B the frontier the node does not replace (sets trade places by value, so the caller's slots
// never move).  Ordinals: row r of the band at bits 2r of po[0] (match) / po[32] (X-gap); the low
// half writes the words, the high half ORs its rows in one iteration later.
struct HalfRef {   // one half of the lane-interleaved words of a frontier set: element k at word [k * 32]
  uint32_t *w;
  uint32_t sh;     // 0 = low half, 16 = high half
  EL_HD uint32_t get(int k) const { return (w[k * 32] >> sh) & 0xffffu; }
  EL_HD void set(int k, uint32_t v) const { w[k * 32] = (w[k * 32] & ~(0xffffu << sh)) | ((v & 0xffffu) << sh); }
};

__host__ __device__ __noinline__ inline uint32_t arrange_half(uint32_t *wa, uint32_t *wb, int R, int row0, int ordrow0, bool low,
                                                              uint32_t ra, int kindA, int kindB, int open, int ext, uint32_t *po) {
  const int m = (ra >> 8) & 3;
  const HalfRef sa{wa, low ? 0u : 16u}, sb{wb, low ? 0u : 16u};
  const int nel = low ? 2 * R + 1 : 2 * R;
  auto swap_sets = [&]() {
    for (int k = 0; k < nel; ++k) { const uint32_t t = sa.get(k); sa.set(k, sb.get(k)); sb.set(k, t); }
    const int k = kindA; kindA = kindB; kindB = k;
  };
  auto vS = [&](int row) { return kBiasP + (row < 0 ? 0 : -(open + ext * row)); };          // virtual column -1
  auto vG = [&](int row) { return kBiasP + (row < 0 ? -open : -(open + ext * row) - ext); };
  if (ra & NF_TWO) {
    if (kindA == 2) swap_sets();   // list order: ref predecessor, then cor
  } else if (!(ra & NF_NOPRED)) {
    const int pk = (ra & NF_PREDC) ? 2 : 1;
    if (!(kindA & pk)) swap_sets();
    if (kindA & ~m) {              // the node leaves one of A's frontiers behind: B := A
      for (int k = 0; k < nel; ++k) sb.set(k, sa.get(k));
      kindB = kindA & ~m;
    }
  } else if (kindA & ~m) swap_sets();
  if (ra & NF_NOPRED) {
    if (low) sa.set(2 * R, (uint32_t)vS(row0 - 1));
    for (int r = 0; r < R; ++r) { sa.set(r, (uint32_t)vS(row0 + r)); sa.set(R + r, (uint32_t)vG(row0 + r)); }
  } else if (ra & (NF_VIRT | NF_TWO)) {
    const bool virt = ra & NF_VIRT, two = ra & NF_TWO;
    const uint32_t oA = virt ? 1u : 0u, oB = oA + 1u;
    uint32_t owM = 0, owX = 0;
    if (low) {
      int bS = (int)sa.get(2 * R); uint32_t o = oA;
      if (virt) { const int a = bS; bS = vS(row0 - 1); o = 0; if (a > bS) { bS = a; o = oA; } }
      if (two) { const int hb = (int)sb.get(2 * R); if (hb > bS) { bS = hb; o = oB; } }
      sa.set(2 * R, (uint32_t)bS); owM |= o;
    }
    for (int r = 0; r < R; ++r) {
      const int aS = (int)sa.get(r), aG = (int)sa.get(R + r);
      int bS = aS, bG = aG; uint32_t oM = oA, oX = oA;
      if (virt) {
        bS = vS(row0 + r); bG = vG(row0 + r); oM = oX = 0;
        if (aS > bS) { bS = aS; oM = oA; }
        if (aG > bG) { bG = aG; oX = oA; }
      }
      if (two) {
        const int sS = (int)sb.get(r), sG = (int)sb.get(R + r);
        if (sS > bS) { bS = sS; oM = oB; }
        if (sG > bG) { bG = sG; oX = oB; }
      }
      sa.set(r, (uint32_t)bS); sa.set(R + r, (uint32_t)bG);
      const int mr = ordrow0 + r + 1;                        // the row whose match move starts from this cell
      if (mr < 2 * R) owM |= oM << (2 * mr);                 // the band's last row feeds the next band's h
      owX |= oX << (2 * (ordrow0 + r));
    }
    if (low) { po[0] = owM; po[32] = owX; }
    else { po[0] |= owM; po[32] |= owX; }
  }
  kindA = m;
  kindB &= ~m;
  return (uint32_t)kindA | ((uint32_t)kindB << 2);
}

struct Phase2P {
  typedef Layout2P Layout;
  static constexpr bool kGenericSub = false;
  static constexpr bool kLinear = false;
  static constexpr bool kBanded = false;
  static constexpr int kSetWords = kSetWordsP;
  static constexpr uint32_t kRecNode = Q2_NODE, kRecX2Y = Q2_X2Y, kRecPred = Q2_PRED;
  static EL_HD void make_layout(Layout2P &L, int N1, int LU) { make_layout2p(L, N1, LU); }
  static EL_HD void put_row0(uint32_t *p, int bS, int bG) { p[Q2_BSG * 32] = pk2(kBiasP + bS, kBiasP + bG); }
  LaneScratch scr;
  uint32_t *bset;     // two packed frontier-set slots of kSetWordsP words, + lane (shared memory on the device)
  Scoring sc;
  const Layout2P *Lp;

  EL_HD uint32_t *rec(int j) const { return scr.at(Lp->o_nodes + (uint32_t)(j + 1) * Lp->rec_words); }
  EL_HD uint32_t *node_rec(int j) const { return rec(j); }

  // one band of 2R rows of DP2 (align_lpo_po2.c:269-433)
  template <int R>
  EL_HDN void band(int nx, int ly, int b, bool last, int &best, int &best_j) const {
    const int r0 = b * 2 * R;
    PackedConsts pc;
    pc.set(sc);
    uint32_t y2[R], S[R], G[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      y2[k] = ((uint32_t)scr.code_at(Lp->o_unc, r0 + k) | ((uint32_t)scr.code_at(Lp->o_unc, r0 + R + k) << 16)) << 4;
      S[k] = G[k] = 0;
    }
    const int rr = ly - 1 - r0;                                // row of the last cell inside this band (when last)
    const uint32_t step = Lp->rec_words * 32;
    uint32_t *p = rec(-1);                                     // record of node j - 1
    uint32_t ra = p[step + Q2_NODE * 32], bsg = p[step + Q2_BSG * 32];   // node j: shape, S | G << 16 of the row above the band
    uint32_t ra_prev = 0, hA = 0, d7 = 0, x2 = 0, mlo = 0;
    uint32_t kinds = 0;   // which frontiers sets A / B hold (1 ref, 2 cor): A low | B low << 2 | A high << 4 | B high << 6
    bool unc_prev = false;
    for (int j = 0; j <= nx; ++j, p += step) {
      // prefetch node j + 1 (the last two iterations re-read node nx - 1: defined values, results unused)
      const uint32_t *pn = p + (j + 1 < nx ? 2 * step : step);
      const uint32_t ra_n = pn[Q2_NODE * 32], bsg_n = pn[Q2_BSG * 32];
      const int m = (ra >> 8) & 3;
      const bool unc_now = j < nx && ((ra & (NF_TWO | NF_NOPRED | NF_VIRT | NF_PREDC)) || (int)(kinds & 3u) != m);
      const uint32_t up0 = (bsg >> 16) | (G[R - 1] << 16);     // G(r0-1, j) | G(r0+R-1, j-1)
      // -- uncommon: bring the node's source column into set A; low half for node j, high half for node j - 1
      if (unc_now || unc_prev) {
        uint32_t *wa = bset, *wb = bset + kSetWordsP;
#pragma unroll
        for (int k = 0; k < R; ++k) { wa[k * 32] = S[k]; wa[(R + k) * 32] = G[k]; }
        wa[2 * R * 32] = hA;
        if (unc_now) {
          uint32_t *po = nullptr;
          if (ra & (NF_VIRT | NF_TWO)) po = scr.at(Lp->o_ord + ((ra >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
          const uint32_t st = arrange_half(wa, wb, R, r0, 0, true, ra, kinds & 3, (kinds >> 2) & 3, sc.open, sc.ext, po);
          kinds = (kinds & ~15u) | st;
        }
        if (unc_prev) {
          uint32_t *po = nullptr;
          if (ra_prev & (NF_VIRT | NF_TWO)) po = scr.at(Lp->o_ord + ((ra_prev >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
          const uint32_t st = arrange_half(wa, wb, R, r0 + R, R, false, ra_prev, (kinds >> 4) & 3, (kinds >> 6) & 3, sc.open, sc.ext, po);
          kinds = (kinds & 15u) | (st << 4);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) { S[k] = wa[k * 32]; G[k] = wa[(R + k) * 32]; }
        hA = wa[2 * R * 32];
      }
      const uint32_t diag0 = (hA & 0xffffu) | (d7 << 16);      // S(r0-1, source of j) | S(r0+R-1, source of j-1)
      d7 = S[R - 1];
      x2 = (x2 << 16) | ((ra & 0xffu) << 4);
      const uint32_t mv = update_packed<R>(pc, S, G, y2, x2, diag0, up0);
      hA = bsg;                                                // the node's own column is now the frontier in set A
      // node j - 1 is now complete in the high half
      if (!last) p[Q2_BSG * 32] = (S[R - 1] >> 16) | (G[R - 1] & 0xffff0000u);
      p[(Q2_MOVES + b) * 32] = pk_lo_hi(mlo, mv);
      mlo = mv;
      if (last) {                                              // best FINAL node; ties keep the smaller j (:410-417)
        if (rr < R) {
          if (j < nx && (ra & NF_FINAL)) { const int s = pick_half<R>(S, rr, false); if (s > best) { best = s; best_j = j; } }
        } else if (j >= 1 && (ra_prev & NF_FINAL)) {
          const int s = pick_half<R>(S, rr - R, true);
          if (s > best) { best = s; best_j = j - 1; }
        }
      }
      ra_prev = ra; unc_prev = unc_now; ra = ra_n; bsg = bsg_n;
    }
  }

  template <int R>
  EL_HDN int dp(int nx, int ly, int &best_j) const {
    const int nb = (ly + 2 * R - 1) / (2 * R);
    int best = -999999;
    best_j = -1;
    for (int b = 0; b < nb - 1; ++b) band<R>(nx, ly, b, false, best, best_j);
    band<R>(nx, ly, nb - 1, true, best, best_j);
    return best;
  }

  // traceback (align_lpo_po2.c:108-168): fills the x2y field of the node records
  template <int R>
  EL_HDN void traceback(int ly, int best_j) const {
    const ptrdiff_t step = (ptrdiff_t)Lp->rec_words * 32;
    int j = best_j, r = ly - 1;
    uint32_t ntw = 0;
    int ntbase = -1;
    while (j >= 0 && r >= 0) {
      const int b = r / (2 * R);
      int rr = r - b * 2 * R;
      uint32_t *p = rec(j);
      const uint32_t *pm = p + (Q2_MOVES + b) * 32;
      uint32_t w0 = pm[0], w1 = j >= 1 ? pm[-step] : 0, w2 = j >= 2 ? pm[-2 * step] : 0, w3 = j >= 3 ? pm[-3 * step] : 0;
      for (;;) {
        const uint32_t kind = (w0 >> (rr < R ? 2 * rr : 16 + 2 * (rr - R))) & 3u;   // bit 1 match, bit 0 X-gap
        if (kind & 2u) p[Q2_X2Y * 32] = (uint32_t)r;
        bool jump = false;
        if (kind) {  // match or X-gap: step to a predecessor of j
          if ((j >> 5) != ntbase) { ntbase = j >> 5; ntw = scr.w(Lp->o_nt + ntbase); }
          if ((ntw >> (j & 31)) & 1u) {
            const uint32_t ra = p[Q2_NODE * 32];
            int ord = 0;
            if (ra & (NF_VIRT | NF_TWO)) {
              const uint32_t *po = scr.at(Lp->o_ord + ((ra >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
              ord = (int)((((kind & 2u) ? po[0] : po[32]) >> (2 * rr)) & 3u);
            }
            const uint32_t pr = p[Q2_PRED * 32];
            const int pA = (pr & 0xffffu) == 0xffffu ? -1 : (int)(pr & 0xffffu), pB = (pr >> 16) == 0xffffu ? -1 : (int)(pr >> 16);
            if (ra & NF_VIRT) j = (ord == 0) ? -1 : (ord == 1 ? pA : pB);
            else j = (ord == 0) ? pA : pB;  // pA == -1 when the list is the virtual link alone
            jump = true;
          } else {
            --j; p -= step; pm -= step;
            w0 = w1; w1 = w2; w2 = w3;
            w3 = j >= 3 ? pm[-3 * step] : 0;
          }
        }
        if (kind != 1u) { --r; --rr; }  // match or Y-gap: step up
        if (jump || j < 0 || r < 0 || rr < 0) break;
      }
    }
  }

  template <int R>
  EL_HDN int run_r(int n1, int lu, int &s2) const {
    int bj;
    s2 = dp<R>(n1, lu, bj);
    traceback<R>(lu, bj);
    return fuse_emit_rows(*this, n1, lu);
  }

  EL_HDN int run_window(const uint16_t *p1, int n1, const uint8_t *unc, int lu, int &s2) const {
    scr.pack_codes(sc.tab, unc, lu, Lp->o_unc);
    for (uint32_t k = 0; k < 5; ++k) scr.w(Lp->o_unc + cdiv_u((uint32_t)lu, 4) + k) = 0;   // rows past the end: defined letters
    prepare_nodes(*this, p1, n1);
    if (Lp->R == 6) return run_r<6>(n1, lu, s2);   // R is uniform over the warp
    if (Lp->R == 7) return run_r<7>(n1, lu, s2);
    return run_r<8>(n1, lu, s2);
  }
};

// =============================== phase 2 of a window whose P1 is linear =========================
// ref and cor identical (spcode 0, most windows at ELECTOR's corrected error rates): P1 = lin(ref) with
// every node carrying both letters, so DP2 is the linear x linear DP of phase 1 with unc as the row
// sequence.  Phase2L runs Phase1P's bands and traceback on (ref, unc) and emits the three MSA rows
// (lpo.c:413-463, lpo_format.c:346-371 for a linear x): no node list, no frontier sets, no ordinals.
struct Layout2L {
  Layout1P dp;                      // o_ref = ref codes (columns), o_cor = unc codes (rows)
  uint32_t o_rows, row_words, total;
};
EL_HD void make_layout2l(Layout2L &L, int N1, int LU) {
  make_layout1p(L.dp, N1, LU);
  uint32_t o = L.dp.total;
  L.row_words = cdiv_u(N1 + LU, 4);
  L.o_rows = o; o += 3 * L.row_words;
  L.total = o;
}

struct Phase2L {
  typedef Layout2L Layout;
  static constexpr bool kGenericSub = false;
  static constexpr bool kLinear = true;
  static constexpr bool kBanded = true;
  static constexpr int kSetWords = 1;
  static EL_HD void make_layout(Layout2L &L, int N1, int LU) { make_layout2l(L, N1, LU); }
  LaneScratch scr;
  uint32_t *bset;     // unused
  Scoring sc;
  const Layout2L *Lp;
  BandW bw = {0, 0, 0, false};   // set per group by the kernel

  // rows of the MSA from the x -> y map of the traceback; returns nring
  EL_HDN int emit(const Phase1P &d, int n1, int lu) const {
    const uint8_t *sym = sc.tab->sym;
    const uint32_t o_x = Lp->dp.o_ref, o_y = Lp->dp.o_cor;
    const uint32_t r0 = Lp->o_rows, r1 = Lp->o_rows + Lp->row_words, r2 = Lp->o_rows + 2 * Lp->row_words;
    int col = 0, iy = 0;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    auto put = [&](uint32_t c0, uint32_t c1, uint32_t c2) {
      const int sh = (col & 3) * 8;
      w0 |= c0 << sh; w1 |= c1 << sh; w2 |= c2 << sh;
      if ((col & 3) == 3) { scr.w(r0 + (col >> 2)) = w0; scr.w(r1 + (col >> 2)) = w1; scr.w(r2 + (col >> 2)) = w2; w0 = w1 = w2 = 0; }
      ++col;
    };
    const ptrdiff_t step = (ptrdiff_t)Lp->dp.rec_words * 32;
    const uint32_t *px = d.rec(0) + P1_X2Y * 32;
    int q0 = (int)px[0], q1 = n1 > 1 ? (int)px[step] : -1;
    uint32_t xw = 0;
    for (int ix = 0; ix < n1; ++ix) {
      const int q = q0;
      q0 = q1;
      q1 = ix + 2 < n1 ? (int)px[(ptrdiff_t)(ix + 2) * step] : -1;
      if ((ix & 3) == 0) xw = scr.w(o_x + (ix >> 2));
      const uint32_t xl = xw & 0xff; xw >>= 8;
      const uint32_t xc = sym[xl];
      if (q >= 0) while (iy < q) { put('.', '.', sym[scr.code_at(o_y, iy)]); ++iy; }   // unaligned unc letters: columns of their own
      uint32_t c2 = '.';
      if (q >= 0 && iy < lu) { c2 = sym[scr.code_at(o_y, iy)]; ++iy; }   // aligned: same column (same letter or same ring)
      put(xc, xc, c2);
    }
    while (iy < lu) { put('.', '.', sym[scr.code_at(o_y, iy)]); ++iy; }
    if (col & 3) { scr.w(r0 + (col >> 2)) = w0; scr.w(r1 + (col >> 2)) = w1; scr.w(r2 + (col >> 2)) = w2; }
    return col;
  }

  template <int R>
  EL_HDN int run_r(const Phase1P &d, int n1, int lu, int &s2, bool &exact) const {
    s2 = d.dp<R>(n1, lu, d.bw);
    exact = d.band_exact(s2, n1, lu);
    if (!exact) return 0;
    d.traceback<R>(n1, lu);
    return emit(d, n1, lu);
  }

  EL_HDN int run_linear(const uint8_t *ref, int n1, const uint8_t *unc, int lu, int &s2, bool &exact) const {
    Phase1P d;
    d.scr = scr; d.sc = sc; d.Lp = &Lp->dp; d.bw = bw;
    scr.pack_codes(sc.tab, ref, n1, Lp->dp.o_ref);
    scr.pack_codes(sc.tab, unc, lu, Lp->dp.o_cor);
    for (uint32_t k = 0; k < 5; ++k) scr.w(Lp->dp.o_cor + cdiv_u((uint32_t)lu, 4) + k) = 0;
    scr.w(Lp->dp.o_ref + cdiv_u((uint32_t)n1, 4)) = 0;
    if (Lp->dp.R == 6) return run_r<6>(d, n1, lu, s2, exact);   // R is uniform over the warp
    if (Lp->dp.R == 7) return run_r<7>(d, n1, lu, s2, exact);
    return run_r<8>(d, n1, lu, s2, exact);
  }
};

}  // namespace elector

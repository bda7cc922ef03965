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
    mis2 = sc.mis2; nopen2 = sc.nopen2; ext2 = sc.ext2;
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
enum : uint32_t { P1_BSG = 0, P1_MOVES = 1 };   // boundary S | G << 16 below the band, one moves word per band

struct Layout1P {
  uint32_t o_nodes, rec_words, R;
  uint32_t o_fast, f_ref, f_cor, f_xb, f_yb, f_total;   // fast part: letter codes and alignment bitmaps (see poa_kernel.cuh)
  uint32_t total;
};
EL_HD void make_layout1p(Layout1P &L, int LR, int LC) {
  uint32_t f = 0;
  L.f_ref = f; f += cdiv_u(LR, 4) + 3;                         // the last iteration reads one letter past the end, the look-ahead one word more
  L.f_cor = f; f += cdiv_u(LC, 4) + 5;                         // rows up to 16 * ceil(LC / 16) - 1 are read
  L.f_xb = f; f += cdiv_u(LR, 32) + 1;
  L.f_yb = f; f += cdiv_u(LC, 32) + 1;
  L.f_total = f;
  uint32_t o = 0;
  L.R = (uint32_t)packed_rows(LC);
  L.rec_words = P1_MOVES + cdiv_u(LC, 16);
  L.o_nodes = o; o += ((uint32_t)LR + 4) * L.rec_words;        // record 0, LR nodes, three records of look-ahead
  L.o_fast = o; o += f;
  L.total = o;
}

struct Phase1P {
  typedef Layout1P Layout;
  static constexpr bool kGenericSub = false;
  static constexpr bool kBanded = true;
  LaneScratch scr;    // slow part of the layout (global scratch)
  LaneScratch fs;     // fast part (shared-memory arena, or global scratch at o_fast)
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
      y2[k] = ((uint32_t)fs.code_at(Lp->f_cor, r0 + k) | ((uint32_t)fs.code_at(Lp->f_cor, r0 + R + k) << 16)) << 4;
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
    uint32_t xw = 0, xw_next = fs.w(Lp->f_ref + (jlo >> 2)), x2 = 0, d7 = 0, mlo = 0;
    if (jlo > 0) { x2 = (uint32_t)fs.code_at(Lp->f_ref, jlo - 1) << 4; d7 = (uint32_t)kNegP; }
    int best = 0;
    for (int j = jlo; j <= jend; ++j, p += step) {             // p = record of node j - 1
      if ((j & 3) == 0 || j == jlo) { xw = xw_next >> (8 * (j & 3)); xw_next = fs.w(Lp->f_ref + (j >> 2) + 1); }   // the next letters, one group ahead
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
      st_stream(p + (P1_MOVES + b) * 32, pk_lo_hi(mlo, mv));
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

  // traceback (align_lpo_po2.c:108-168): marks the aligned pairs in the bitmaps.  One uniform step per cell of the path
  // (load the moves word of (node, band), decode, mark, move): the lanes of a warp stay in lock step whatever their paths
  // do -- a walk with a register queue of moves words and an inner loop per band executed ~4x the instructions at warp
  // level (profiles/r2b: divergence, not work).  The lines of the cells ahead are requested into L1 as the walk goes.
  template <int R>
  EL_HDN void traceback(int lr, int ly, AlignBits &al) const {
    al.clear(lr, ly);
    const uint32_t *base = rec(0) + P1_MOVES * 32;             // moves word of (node j, band b) at base[(j * rec_words + b) * 32]
    const int rw = (int)Lp->rec_words;
    int j = lr - 1, r = ly - 1;
    int b = r / (2 * R), rr = r - b * 2 * R;
    int nmatch = 0;
    while (j >= 0 && r >= 0) {
      const uint32_t *pm = base + (ptrdiff_t)(j * rw + b) * 32;
      const uint32_t w = *pm;
      if (j >= 6) prefetch_l1(pm - (ptrdiff_t)6 * rw * 32);
      if (b > 0 && j >= 2) prefetch_l1(pm - (ptrdiff_t)2 * rw * 32 - 32);
      const uint32_t kind = (w >> (rr < R ? 2 * rr : 16 + 2 * (rr - R))) & 3u;   // bit 1 match, bit 0 X-gap
      if (kind & 2u) {
        al.st.w(al.ox + (uint32_t)(j >> 5)) |= 1u << (j & 31);
        al.st.w(al.oy + (uint32_t)(r >> 5)) |= 1u << (r & 31);
        ++nmatch;
      }
      const int dr = kind != 1u, dj = kind != 0u;
      j -= dj; r -= dr; rr -= dr;
      if (rr < 0) { rr += 2 * R; --b; }
    }
    al.nmatch = nmatch;
  }

  template <int R>
  EL_HDN int run_r(int lr, int lc, uint16_t *p1_out, int &s1, int &spcode, bool &exact) const {
    s1 = dp<R>(lr, lc, bw);
    exact = band_exact(s1, lr, lc);
    if (!exact) return 0;                                      // to be run again without a band
    AlignBits al{fs, Lp->f_xb, Lp->f_yb};
    traceback<R>(lr, lc, al);
    return fuse1(fs, Lp->f_ref, Lp->f_cor, al, lr, lc, p1_out, spcode);
  }

  // the codes of both sequences, with defined values where the band loops read past the ends (up to 15 letters past the
  // end of the row sequence, one past the end of the column sequence; which values does not matter: they only feed cells
  // outside the window)
  EL_HDN void pack(const uint8_t *x, int lx, const uint8_t *y, int ly) const {
    fs.pack_codes(sc.tab, x, lx, Lp->f_ref);
    fs.pack_codes(sc.tab, y, ly, Lp->f_cor);
    for (uint32_t k = 0; k < 5; ++k) fs.w(Lp->f_cor + cdiv_u((uint32_t)ly, 4) + k) = 0;
    fs.w(Lp->f_ref + cdiv_u((uint32_t)lx, 4)) = 0;
    fs.w(Lp->f_ref + cdiv_u((uint32_t)lx, 4) + 1) = 0;
  }

  // exact = false (band on only): nothing was written to p1_out, the window has to be run again without a band
  EL_HDN int run_window(const uint8_t *ref, int lr, const uint8_t *cor, int lc, uint16_t *p1_out, int &s1, int &spcode, bool &exact) const {
    pack(ref, lr, cor, lc);
    if (Lp->R == 6) return run_r<6>(lr, lc, p1_out, s1, spcode, exact);   // R is uniform over the warp
    if (Lp->R == 7) return run_r<7>(lr, lc, p1_out, s1, spcode, exact);
    return run_r<8>(lr, lc, p1_out, s1, spcode, exact);
  }
};

// =============================== phase 2 of a window whose P1 is linear =========================
// ref and cor identical (spcode 0, most windows at ELECTOR's corrected error rates): P1 = lin(ref) with
// every node carrying both letters, so DP2 is the linear x linear DP of phase 1 with unc as the row
// sequence.  Phase2L runs Phase1P's bands and traceback on (ref, unc) and emits the three MSA rows
// (lpo.c:413-463, lpo_format.c:346-371 for a linear x): no node list, no frontier sets, no ordinals.
// (General windows run the dual-frontier kernel of poa_dual.cuh.)
struct Layout2L : Layout1P {};         // f_ref = ref codes (columns), f_cor = unc codes (rows)
EL_HD void make_layout2l(Layout2L &L, int N1, int LU) { make_layout1p(L, N1, LU); }

struct Phase2L {
  typedef Layout2L Layout;
  static constexpr bool kGenericSub = false;
  static constexpr bool kLinear = true;
  static constexpr bool kBanded = true;
  static constexpr int kSetWords = 1;
  static EL_HD void make_layout(Layout2L &L, int N1, int LU) { make_layout2l(L, N1, LU); }
  LaneScratch scr, fs;
  uint32_t *bset;     // unused
  Scoring sc;
  const Layout2L *Lp;
  BandW bw = {0, 0, 0, false};   // set per group by the kernel
#ifdef EL_DP_CLOCKS
  PhaseClock pclk;
#endif

  EL_HD AlignBits bits() const { return AlignBits{fs, Lp->f_xb, Lp->f_yb}; }
  EL_HD Phase1P core() const {
    Phase1P d;
    d.scr = scr; d.fs = fs; d.sc = sc; d.Lp = Lp; d.bw = bw;
    return d;
  }

  // rows of the MSA from the alignment bitmaps of the traceback, straight to their place; returns nring.
  // Column-driven and branch-free: a column is a node of lin(ref) (with the letter of unc aligned to it, if any) or an
  // unaligned letter of unc, which goes before the next ALIGNED node or after the last node (lpo.c:413-463 for a linear x).
  // The corrected row of these windows is the reference row.
  EL_HDN int fuse_emit(const AlignBits &al, int n1, int lu, const RowSink &out) const {
    const uint8_t *sym = sc.tab->sym;
    int col = 0, ix = 0, iy = 0;
    uint32_t w0 = 0, w2 = 0;
    while (ix < n1 || iy < lu) {
      const bool xa = ix < n1 && ((fs.w(al.ox + (uint32_t)(ix >> 5)) >> (ix & 31)) & 1u);
      const bool ya = iy < lu && ((fs.w(al.oy + (uint32_t)(iy >> 5)) >> (iy & 31)) & 1u);
      const bool yonly = iy < lu && !ya && (ix >= n1 || xa);
      const bool takey = yonly || xa;                          // an aligned node's partner is the current letter of unc
      const uint32_t xc = sym[(fs.w(Lp->f_ref + (uint32_t)(ix >> 2)) >> ((ix & 3) * 8)) & 31u];
      const uint32_t yc = sym[(fs.w(Lp->f_cor + (uint32_t)(iy >> 2)) >> ((iy & 3) * 8)) & 31u];
      const int sh = (col & 3) * 8;
      w0 |= (yonly ? (uint32_t)'.' : xc) << sh;
      w2 |= (takey ? yc : (uint32_t)'.') << sh;
      if ((col & 3) == 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w0); st_stream(out.r2 + (col >> 2), w2); w0 = w2 = 0; }
      ++col;
      ix += !yonly; iy += takey;
    }
    if (col & 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w0); st_stream(out.r2 + (col >> 2), w2); }
    return col;
  }

  template <int R>
  EL_HDN int run_r(const Phase1P &d, int n1, int lu, int &s2, bool &exact, AlignBits &al) const {
    s2 = d.dp<R>(n1, lu, d.bw);
    EL_TICK(*this, 3);
    exact = d.band_exact(s2, n1, lu);
    if (!exact) return 0;
    d.traceback<R>(n1, lu, al);
    EL_TICK(*this, 4);
    return columns_of(n1, lu, al.nmatch);   // lin(ref): every node is a ring of its own
  }

  // everything up to the traceback; returns the number of MSA columns (0 and exact = false: run again without the band)
  EL_HDN int align_linear(const uint8_t *ref, int n1, const uint8_t *unc, int lu, int &s2, bool &exact, AlignBits &al) const {
    const Phase1P d = core();
    d.pack(ref, n1, unc, lu);
    EL_TICK(*this, 1);
    if (Lp->R == 6) return run_r<6>(d, n1, lu, s2, exact, al);   // R is uniform over the warp
    if (Lp->R == 7) return run_r<7>(d, n1, lu, s2, exact, al);
    return run_r<8>(d, n1, lu, s2, exact, al);
  }
};

}  // namespace elector

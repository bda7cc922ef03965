// poa_dual.cuh -- DP2 of the GENERAL windows (ref and cor differ) without frontier-set juggling, two cells per instruction.
//
// P1 is the partial order of two sequences: every node has its predecessor on the ref path (the latest
// ref-carrying node), on the cor path (the latest cor-carrying node), or both.  Phase2 (poa_kernel.cuh)
// keeps ONE frontier column in registers and a spare in shared memory and swaps / copies / merges the two
// out of line whenever a node needs the other frontier -- the whole warp waits each time one lane does
// (profiles/r1c_ncu_dp2_int32_by_function.txt: 43 % of the kernel's instructions, the column update 37 %).
//
// Here both frontiers live in registers all the time, as two register sets: Sr/Gr hold the column of the latest
// ref-carrying node, Sc/Gc the column of the latest cor-carrying node (both start as the virtual column -1,
// align_lpo_po2.c:272-302).  The two 16-bit halves of a register are two ROWS of that column, skewed like the bands of
// poa_packed.cuh: a band of 2R rows, the low half holds row r0 + k and works on node j, the high half holds row
// r0 + R + k and works on node j - 1 (its first row takes "up" and "diagonal" from the low half's last row of the
// iteration before).  A node then is one straight-line packed update, 2 cells per instruction:
//   * the input column is picked per half by a bit-select (the ref frontier for a node carrying a ref letter, else the cor
//     one), the result is written back, again by bit-selects, to the frontier(s) of the letters the node carries: a node
//     with both letters replaces both, a one-letter node replaces its own and keeps the other -- no data moves, no branch;
//   * only a both-node that FOLLOWS a one-letter node (the end of a bubble) first takes the maximum of the two frontiers:
//     the first strict maximum over its left list (align_lpo_po2.c:334-371), which is [ref predecessor, cor predecessor],
//     or [virtual -1, the one real predecessor] for an INITIAL node -- the frontier that is still the virtual column then
//     plays the virtual link.  Two VIMNMX.S16x2 (a, b) / (b, a) yield the maximum and both strict comparisons as
//     predicates; the winning predecessor ordinals go to the ordinal words the traceback of Phase2 reads.  The low half
//     of a node merges in iteration j, its high half in iteration j + 1 (the ordinals of the low rows wait in registers).
// (Round 1 kept the two frontiers in the two HALVES of one register set: one cell per instruction, 13 instructions per
// cell against 17 per pair of cells here.)
// Arithmetic, bias and matrix class are those of poa_packed.cuh (exact for packed_ok() matrices and scores
// that stay inside 16 bits; everything else runs Phase2).  Node records (the boundary row S | G << 16 in ONE word), moves
// words, ordinals, traceback, fuse and emit are Phase2's: only the band sweep differs.
#pragma once
#include "poa_kernel.cuh"
#include "poa_packed.cuh"

namespace elector {

EL_HD uint32_t pk_select(uint32_t fresh, uint32_t old, uint32_t take) { return (fresh & take) | (old & ~take); }   // one LOP3
// PRMT: result byte i = byte (sel >> 4i & 7) of {a: 0-3, b: 4-7}, or that byte's sign bit replicated when bit 3 of the nibble is set
EL_HD uint32_t pk_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));   // (__byte_perm drops the replicate bit)
  return r;
#else
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 0xfu;
    uint32_t byte = (uint32_t)(ab >> (8 * (n & 7u))) & 0xffu;
    if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
    r |= byte << (8 * i);
  }
  return r;
#endif
}

// shape word of a node (record field R2_BG, free in this kernel), made by the node preparation so that the band loop
// derives everything it needs per node with one PRMT each: bit 7 = carries a ref letter, bit 15 = carries a cor letter,
// bit 23 = closes a bubble (a both-node after a one-letter node: exactly the nodes with an ordinal slot, NF_VIRT | NF_TWO),
// byte 3 = the letter code, bit 0 = NF_FINAL, bit 1 = a bubble-closing INITIAL node whose real predecessor is the ref one
// (the virtual link, played by the cor frontier, comes first in its left list)
enum : uint32_t { NM_REF = 1u << 7, NM_COR = 1u << 15, NM_MERGE = 1u << 23, NM_FINAL = 1u, NM_REFWINS = 2u };

struct Phase2D : Phase2<false> {
  static constexpr int kSetWords = 1;   // no frontier sets in shared memory
  // boundary rows live in the node records in packed form: biased S in the low half, biased G in the high half of R2_BS
  static EL_HD void put_row0(uint32_t *p, int bS, int bG) { p[R2_BS * 32] = pk2(kBiasP + bS, kBiasP + bG); }
  static EL_HD void put_shape(uint32_t *p, uint32_t ra) {
    p[R2_BG * 32] = ((ra & NF_REF) ? NM_REF : 0u) | ((ra & NF_COR) ? NM_COR : 0u) | ((ra & (NF_VIRT | NF_TWO)) ? NM_MERGE : 0u) |
                    ((ra & NF_FINAL) ? NM_FINAL : 0u) | (((ra & NF_VIRT) && !(ra & NF_PREDC)) ? NM_REFWINS : 0u) | ((ra & 0xffu) << 24);
  }
  EL_HDN int prepare(const uint16_t *nodes, int nx) const { return prepare_nodes(*this, nodes, nx); }

  // the ref set := per-half maximum of the two frontiers in the halves of `take`; ORs a 1 into the ordinal words where the
  // SECOND entry of the left list wins (strictly): the cor frontier, or the ref frontier when REF_SECOND
  template <int R, bool REF_SECOND>
  static EL_HD void merge(uint32_t (&Sr)[R], uint32_t (&Gr)[R], const uint32_t (&Sc)[R], const uint32_t (&Gc)[R], uint32_t &hr, uint32_t hc,
                          uint32_t take, uint32_t &oM, uint32_t &oX) {
    const uint32_t hm = REF_SECOND ? pk_maxs_flag(hc, hr, oM, 1u, 0u) : pk_maxs_flag(hr, hc, oM, 1u, 0u);
    if (take & 0xffffu) hr = hm;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const uint32_t bMl = 4u << (2 * k), bMh = R + k + 1 < kBand ? 4u << (2 * (R + k)) : 0u;   // the match of the NEXT row starts from this cell
      const uint32_t bXl = 1u << (2 * k), bXh = 1u << (2 * (R + k));
      const uint32_t ms = REF_SECOND ? pk_maxs_flag(Sc[k], Sr[k], oM, bMl, bMh) : pk_maxs_flag(Sr[k], Sc[k], oM, bMl, bMh);
      const uint32_t mx = REF_SECOND ? pk_maxs_flag(Gc[k], Gr[k], oX, bXl, bXh) : pk_maxs_flag(Gr[k], Gc[k], oX, bXl, bXh);
      Sr[k] = pk_select(ms, Sr[k], take);
      Gr[k] = pk_select(mx, Gr[k], take);
    }
  }

  // one band of 2R rows of DP2 (align_lpo_po2.c:269-433).  Moves words: row rr of the band at bits 2 * (15 - rr) (+1: match);
  // ordinal words: the match ordinal of row rr at bits 2 * rr of po[0] (the S cell of row rr - 1 decides it), the X-gap
  // ordinal of row rr at bits 2 * rr of po[32] -- the formats Phase2::traceback reads.
  template <int R>
  EL_HDN void band(int nx, int ly, int b, bool last, int &best, int &best_j) const {
    constexpr uint32_t kLowMoves = ~0u << (32 - 2 * R);        // moves bits of the low half's rows (0 .. R-1)
    constexpr uint32_t kLowOrdM = (4u << (2 * R)) - 1u;        // match ordinals decided by the low half: rows 0 .. R
    constexpr uint32_t kLowOrdX = (1u << (2 * R)) - 1u;        // X-gap ordinals of rows 0 .. R-1
    const int r0 = b * kBand;
    PackedConsts pc;
    pc.set(sc);
    uint32_t y2[R], Sr[R], Gr[R], Sc[R], Gc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      y2[k] = ((uint32_t)fs.code_at(Lp->f_unc, r0 + k) | ((uint32_t)fs.code_at(Lp->f_unc, r0 + R + k) << 16)) * 0x0101u;   // the code in both bytes of its half
      Sr[k] = Sc[k] = pk2(kBiasP + sc.virt_S(r0 + k), kBiasP + sc.virt_S(r0 + R + k));   // both frontiers start as the virtual column -1
      Gr[k] = Gc[k] = pk2(kBiasP + sc.virt_G(r0 + k), kBiasP + sc.virt_G(r0 + R + k));
    }
    uint32_t hr = (uint32_t)(kBiasP + sc.virt_S(r0 - 1)), hc = hr;   // S of the row above the band, per frontier (low halves)
    const int rr = ly - 1 - r0;                                // row of the final cells inside this band (when last)
    const bool lowrow = rr < R;
    uint32_t *p = rec(0);                                      // record of node j
    const int32_t step = (int32_t)(Lp->rec_words * 32);
    const int32_t o_bs = (int32_t)(R2_BS * 32) - step, o_mv = (int32_t)((R2_MOVES + (uint32_t)b) * 32) - step;   // fields of node j - 1
    // per half (low: node j, high: node j - 1): mr = the node carries a ref letter (input and output frontier), mc = it
    // carries a cor letter, mg = it closes a bubble; x2 = the nodes' letters (like y2); nm_h = shape of node j - 1
    uint32_t mr = 0, mc = 0, mg = 0, x2 = 0, nm_h = 0;
    uint32_t d7 = 0, g7 = 0, mlo = 0, ordM_lo = 0, ordX_lo = 0;
    // node j + 2 is loaded in iteration j (the scratch of a launch exceeds the L2; the long-scoreboard stall was this loop's
    // top stall).  The loads run up to three records past the last node: still this lane's scratch (make_layout2 keeps one
    // spare record; the ordinal words, the letter codes and the fast part behind it are longer than the other two), values never used.
    uint32_t nm = p[R2_BG * 32], bsg = p[R2_BS * 32];
    uint32_t nm_n = p[step + R2_BG * 32], bsg_n = p[step + R2_BS * 32];
    for (int j = 0; j <= nx; ++j, p += step) {                 // the last iteration only completes the high half
      const uint32_t nm_n2 = p[2 * step + R2_BG * 32], bsg_n2 = p[2 * step + R2_BS * 32];
      if (j >= nx) nm = 0;
      mr = pk_prmt(mr, nm, 0x10ccu);                           // low half := bit 7 of nm replicated, high half := the old low half
      mc = pk_prmt(mc, nm, 0x10ddu);
      mg = pk_prmt(mg, nm, 0x10eeu);
      x2 = pk_prmt(x2, nm, 0x1077u);
      if (mg) {
        // end of a bubble in one of the halves: first strict maximum over [ref frontier, cor frontier]; for an INITIAL node
        // whose real predecessor is the ref one the virtual link (the cor frontier) comes first in the list, so the ref
        // frontier has to win strictly.  Only the ref set takes the maximum: the node carries both letters, reads its
        // input there and replaces both frontiers.
        const bool low = mg & 0xffffu;                         // the merging node is node j (a node never merges in both halves at once)
        const uint32_t ram = low ? 0u : (p - step)[R2_NODE * 32];   // node j - 1: its ordinal slot (loaded ahead of the merge, used after it)
        uint32_t oM = 0, oX = 0;                               // ordinal 1 = the second entry of the list won
        if ((low ? nm : nm_h) & NM_REFWINS) merge<R, true>(Sr, Gr, Sc, Gc, hr, hc, mg, oM, oX);
        else merge<R, false>(Sr, Gr, Sc, Gc, hr, hc, mg, oM, oX);
        if (mg & 0xffffu) { ordM_lo = oM & kLowOrdM; ordX_lo = oX & kLowOrdX; }
        else {
          uint32_t *po = scr.at(Lp->o_ord + ((ram >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
          po[0] = ordM_lo | (oM & ~kLowOrdM);
          po[32] = ordX_lo | (oX & ~kLowOrdX);
        }
      }
      // the packed update: low halves on node j, high halves on node j - 1
      const uint32_t diag0 = pk_prmt(pk_select(hr, hc, mr), d7, 0x5410u);   // S(r0-1, pred j) | S(r0+R-1, pred (j-1))
      const uint32_t up0 = pk_prmt(bsg, g7, 0x5432u);                       // G(r0-1, j)      | G(r0+R-1, j-1)
      d7 = pk_select(Sr[R - 1], Sc[R - 1], mr);
      uint32_t mv = 0, diag = diag0, up = up0, s = 0, g = 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const uint32_t pS = pk_select(Sr[k], Sc[k], mr), pG = pk_select(Gr[k], Gc[k], mr);
        const uint32_t t = pk_minu(x2 ^ y2[k], pc.mis2);
        const uint32_t M = diag - t;
        const uint32_t gap = pk_maxs_flag(up, pG, mv, 1u << (2 * (kBand - 1 - k)), 1u << (2 * (kBand - 1 - R - k)));   // X-gap only when it beats the Y-gap
        s = pk_maxs_flag(gap, M, mv, 2u << (2 * (kBand - 1 - k)), 2u << (2 * (kBand - 1 - R - k)));                 // match only when it beats both
        g = pk_addmaxs(M, pc.nopen2, gap - pc.ext2);
        Sr[k] = pk_select(s, Sr[k], mr); Sc[k] = pk_select(s, Sc[k], mc);
        Gr[k] = pk_select(g, Gr[k], mr); Gc[k] = pk_select(g, Gc[k], mc);
        diag = pS; up = g;
      }
      g7 = g;
      hr = pk_select(bsg, hr, mr);                             // (only the low halves of hr / hc are read)
      hc = pk_select(bsg, hc, mc);
      // node j - 1 is now complete in the high half (s, g: its last row)
      if (j > 0) {
        if (!last) p[o_bs] = pk_prmt(s, g, 0x7632u);
        st_stream(p + o_mv, (mlo & kLowMoves) | (mv & ~kLowMoves));
      }
      mlo = mv;
      if (last && ((lowrow ? nm : nm_h) & NM_FINAL)) {
        const int kk = lowrow ? rr : rr - R;
        uint32_t v = pk_select(Sr[0], Sc[0], mr);               // the node's own column
#pragma unroll
        for (int k = 1; k < R; ++k) if (kk == k) v = pk_select(Sr[k], Sc[k], mr);
        const int sf = (int)(lowrow ? v & 0xffffu : v >> 16) - kBiasP;
        if (sf > best) { best = sf; best_j = lowrow ? j : j - 1; }   // ties keep the smaller j (align_lpo_po2.c:410-417)
      }
      nm_h = nm;
      nm = nm_n; bsg = bsg_n;
      nm_n = nm_n2; bsg_n = bsg_n2;
    }
  }

  // fuse 2 + MSA emit for a P1 made of TWO sequences (lpo.c:413-463, lpo_format.c:346-371): an align ring of P1 has one node
  // (both letters, or one) or two (a ref-only and a cor-only node: a substitution), so a ring is one MSA column holding ref's
  // and cor's letter, and the walk is column-driven and branch-free like Phase2L's: per step one ring of P1 (with the letter of
  // unc aligned to one of its nodes, if any) or one unaligned letter of unc, which goes before the next ring that has an
  // ALIGNED node, or after the last node.  The nodes are read from P1's compact 16-bit list (a line or two per window) instead
  // of one scratch record per node (the generic fuse_emit_rows: 170 warp instructions per column and this kernel's top stall,
  // profiles/r3b).  Returns the number of columns.
  mutable const uint16_t *p1n = nullptr;
  EL_HDN int fuse_emit(const AlignBits &al, int n1, int lu, const RowSink &out) const {
    const uint8_t *sym = sc.tab->sym;
    int col = 0, ix = 0, iy = 0;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    uint32_t ra0 = n1 > 0 ? p1n[0] : 0u, ra1 = n1 > 1 ? p1n[1] : 0u, ra2 = n1 > 2 ? p1n[2] : 0u, ra3 = n1 > 3 ? p1n[3] : 0u;   // nodes ix .. ix + 3
    while (ix < n1 || iy < lu) {
      const bool two = ix + 1 < n1 && (ra1 & NF_SAMERING);
      const bool xa = ix < n1 && (al.x_at(ix) || (two && al.x_at(ix + 1)));
      const bool ya = iy < lu && ((fs.w(al.oy + (uint32_t)(iy >> 5)) >> (iy & 31)) & 1u);
      const bool yonly = iy < lu && !ya && (ix >= n1 || xa);
      const bool takey = yonly || xa;                          // an aligned ring's partner is the current letter of unc
      const uint32_t rb = two ? (ra0 | ra1) : ra0;             // which letters the ring carries
      const uint32_t rl = (ra0 & NF_REF) ? ra0 : ra1, cl = (ra0 & NF_COR) ? ra0 : ra1;
      const uint32_t rc = (rb & NF_REF) ? (uint32_t)sym[rl & 31u] : (uint32_t)'.';
      const uint32_t cc = (rb & NF_COR) ? (uint32_t)sym[cl & 31u] : (uint32_t)'.';
      const uint32_t yc = sym[(fs.w(Lp->f_unc + (uint32_t)(iy >> 2)) >> ((iy & 3) * 8)) & 31u];
      const int sh = (col & 3) * 8;
      w0 |= (yonly ? (uint32_t)'.' : rc) << sh;
      w1 |= (yonly ? (uint32_t)'.' : cc) << sh;
      w2 |= (takey ? yc : (uint32_t)'.') << sh;
      if ((col & 3) == 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w1); st_stream(out.r2 + (col >> 2), w2); w0 = w1 = w2 = 0; }
      ++col;
      iy += takey;
      if (!yonly) {
        const int adv = two ? 2 : 1;
        ix += adv;
        if (two) { ra0 = ra2; ra1 = ra3; ra2 = ix + 2 < n1 ? p1n[ix + 2] : 0u; ra3 = ix + 3 < n1 ? p1n[ix + 3] : 0u; }
        else { ra0 = ra1; ra1 = ra2; ra2 = ra3; ra3 = ix + 3 < n1 ? p1n[ix + 3] : 0u; }
      }
    }
    if (col & 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w1); st_stream(out.r2 + (col >> 2), w2); }
    return col;
  }

  EL_HDN int dp(int nx, int ly, int &best_j) const {
    const int nb = (ly + kBand - 1) / kBand;
    int best = -999999;
    best_j = -1;
    for (int b = 0; b < nb - 1; ++b) band<8>(nx, ly, b, false, best, best_j);
    if (ly - (nb - 1) * kBand <= 8) band<4>(nx, ly, nb - 1, true, best, best_j);
    else band<8>(nx, ly, nb - 1, true, best, best_j);
    return best;
  }

  // everything up to the traceback; returns the number of MSA columns
  EL_HDN int align_window(const uint16_t *p1, int n1, const uint8_t *unc, int lu, int &s2, AlignBits &al) const {
    p1n = p1;
    for (int k = 0; k < n1; k += 64) prefetch_l1(p1 + k);      // the node list, for the node preparation (and the fusion)
    fs.pack_codes(sc.tab, unc, lu, Lp->f_unc);
    EL_TICK(*this, 1);
    const int nrings = prepare(p1, n1);
    EL_TICK(*this, 2);
    int bj;
    s2 = dp(n1, lu, bj);
    EL_TICK(*this, 3);
    traceback(n1, lu, bj, al);
    EL_TICK(*this, 4);
    return columns_of(nrings, lu, al.nmatch);
  }
};

}  // namespace elector

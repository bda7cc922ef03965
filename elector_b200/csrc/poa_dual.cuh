// poa_dual.cuh -- DP2 of the GENERAL windows (ref and cor differ) without frontier-set juggling.
//
// P1 is the partial order of two sequences: every node has its predecessor on the ref path (the latest
// ref-carrying node), on the cor path (the latest cor-carrying node), or both.  Phase2 (poa_kernel.cuh)
// keeps ONE frontier column in registers and a spare in shared memory and swaps / copies / merges the two
// out of line whenever a node needs the other frontier -- the whole warp waits each time one lane does
// (profiles/r1c_ncu_dp2_int32_by_function.txt: 43 % of the kernel's instructions, the column update 37 %).
//
// Here the two halves of every 16-bit packed register ARE the two frontiers: the low half holds the column
// of the latest ref-carrying node, the high half the column of the latest cor-carrying node (both start as
// the virtual column -1, align_lpo_po2.c:272-302).  A node then is one straight-line packed update:
//   * a node carrying both letters updates both halves (after a merge, below, they are equal and stay so);
//   * a ref-only node updates the low half and keeps the high half (one bit-select per register), a cor-only
//     node the other way round -- no data moves, no branch;
//   * only a both-node that FOLLOWS a one-letter node (the end of a bubble) takes the maximum of the two
//     halves first: the first strict maximum over its left list (align_lpo_po2.c:334-371), which is
//     [ref predecessor, cor predecessor], or [virtual -1, the one real predecessor] for an INITIAL node --
//     the half that is still the virtual column then plays the virtual link.  One VIMNMX.S16x2 on (x, x with
//     swapped halves) yields the maximum in both halves and both strict comparisons as predicates; the
//     winning predecessor ordinals go to the ordinal words the traceback of Phase2 reads.
// Arithmetic, bias and matrix class are those of poa_packed.cuh (exact for packed_ok() matrices and scores
// that stay inside 16 bits; everything else runs Phase2).  Node records, moves words, ordinals, traceback,
// fuse and emit are Phase2's: only the band sweep differs.
#pragma once
#include "poa_kernel.cuh"
#include "poa_packed.cuh"

namespace elector {

// per-half signed max(a, b); ORs bit into mv_lo / mv_hi where b beats a (b > a) in the low / high half
EL_HD uint32_t pk_maxs_flag2(uint32_t a, uint32_t b, uint32_t &mv_lo, uint32_t &mv_hi, uint32_t bit) {
#ifdef __CUDA_ARCH__
  uint32_t val;
  asm("{.reg .pred pu, pv;\n\t"
      ".reg .s16 rs0, rs1, rs2, rs3;\n\t"
      "max.s16x2 %0, %3, %4;\n\t"
      "mov.b32 {rs0, rs1}, %0;\n\t"
      "mov.b32 {rs2, rs3}, %3;\n\t"
      "setp.eq.s16 pv, rs0, rs2;\n\t"
      "setp.eq.s16 pu, rs1, rs3;\n\t"
      "@!pv or.b32 %1, %1, %5;\n\t"
      "@!pu or.b32 %2, %2, %5;}\n\t"
      : "=&r"(val), "+r"(mv_lo), "+r"(mv_hi) : "r"(a), "r"(b), "r"(bit));   // early clobber: a is read after val is written
  return val;
#else
  const int16_t al = (int16_t)(a & 0xffffu), bl = (int16_t)(b & 0xffffu), ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
  if (bl > al) mv_lo |= bit;
  if (bh > ah) mv_hi |= bit;
  return (uint32_t)(uint16_t)(al >= bl ? al : bl) | ((uint32_t)(uint16_t)(ah >= bh ? ah : bh) << 16);
#endif
}
EL_HD uint32_t pk_swap(uint32_t x) { return (x >> 16) | (x << 16); }
EL_HD uint32_t pk_both(uint32_t half) { return (half & 0xffffu) * 0x00010001u; }          // a 16-bit value in both halves
EL_HD uint32_t pk_select(uint32_t fresh, uint32_t old, uint32_t keep) { return (fresh & ~keep) | (old & keep); }   // one LOP3

struct Phase2D : Phase2<false> {
  static constexpr int kSetWords = 1;   // no frontier sets in shared memory
  // boundary rows live in the node records in packed form: the biased value in both halves
  static EL_HD void put_row0(uint32_t *p, int bS, int bG) { p[R2_BS * 32] = pk_both((uint32_t)(kBiasP + bS)); p[R2_BG * 32] = pk_both((uint32_t)(kBiasP + bG)); }
  EL_HDN int prepare(const uint16_t *nodes, int nx) const { return prepare_nodes(*this, nodes, nx); }

  // one band of R rows of DP2 (align_lpo_po2.c:269-433)
  template <int R>
  EL_HDN void band(int nx, int ly, int b, bool last, int &best, int &best_j) const {
    const int r0 = b * kBand;
    PackedConsts pc;
    pc.set(sc);
    uint32_t y2[R], S[R], G[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      y2[r] = pk_both((uint32_t)fs.code_at(Lp->f_unc, r0 + r)) << 4;
      S[r] = pk_both((uint32_t)(kBiasP + sc.virt_S(r0 + r)));       // both frontiers start as the virtual column -1
      G[r] = pk_both((uint32_t)(kBiasP + sc.virt_G(r0 + r)));
    }
    uint32_t h = pk_both((uint32_t)(kBiasP + sc.virt_S(r0 - 1)));   // S of the row above the band, per frontier
    bool synced = true;                                             // both halves hold the same column
    const int rr = ly - 1 - r0;
    uint32_t *p = rec(0);
    const uint32_t step = Lp->rec_words * 32;
    // node j + 2 is loaded in iteration j (the scratch of a launch exceeds the L2; the long-scoreboard stall was this loop's top stall)
    uint32_t ra = p[R2_NODE * 32], bs = p[R2_BS * 32], bg = p[R2_BG * 32];
    const uint32_t *p1 = nx > 1 ? p + step : p;
    uint32_t ra_n = p1[R2_NODE * 32], bs_n = p1[R2_BS * 32], bg_n = p1[R2_BG * 32];
    for (int j = 0; j < nx; ++j, p += step) {
      const uint32_t *pn = j + 2 < nx ? p + 2 * step : p;
      const uint32_t ra_n2 = pn[R2_NODE * 32], bs_n2 = pn[R2_BS * 32], bg_n2 = pn[R2_BG * 32];
      const bool has_r = ra & NF_REF, both = has_r && (ra & NF_COR);
      if (both && !synced) {
        // end of a bubble: first strict maximum over [low half, high half]; for an INITIAL node whose real predecessor is
        // the ref one the virtual link (the high half) comes first in the list, so the low half has to win strictly
        uint32_t gtM = 0, ltM = 0, gtX = 0, ltX = 0;
        h = pk_maxs_flag2(h, pk_swap(h), gtM, ltM, 1u);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r + 1 < kBand) S[r] = pk_maxs_flag2(S[r], pk_swap(S[r]), gtM, ltM, 1u << (2 * (r + 1)));   // row r+1's match starts here
          else { uint32_t d0 = 0, d1 = 0; S[r] = pk_maxs_flag2(S[r], pk_swap(S[r]), d0, d1, 0u); }       // the next band's halo
          G[r] = pk_maxs_flag2(G[r], pk_swap(G[r]), gtX, ltX, 1u << (2 * r));
        }
        const bool low_must_win = (ra & NF_VIRT) && !(ra & NF_PREDC);
        uint32_t *po = scr.at(Lp->o_ord + ((ra >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
        po[0] = low_must_win ? ltM : gtM;
        po[32] = low_must_win ? ltX : gtX;
      }
      synced = both;
      const uint32_t keep = both ? 0u : has_r ? 0xffff0000u : 0x0000ffffu;   // the half this node does not replace
      const uint32_t x2 = pk_both(ra & 0xffu) << 4;
      uint32_t mvl = 0, mvh = 0;
      uint32_t diag = h, up = bg;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t pS = S[r], pG = G[r];
        const uint32_t t = pk_minu(x2 ^ y2[r], pc.mis2);
        const uint32_t M = diag - t;
        const uint32_t gap = pk_maxs_flag2(up, pG, mvl, mvh, 1u << (2 * (kBand - 1 - r)));        // X-gap only when it beats the Y-gap
        const uint32_t s = pk_maxs_flag2(gap, M, mvl, mvh, 2u << (2 * (kBand - 1 - r)));          // match only when it beats both
        const uint32_t g = pk_addmaxs(M, pc.nopen2, gap - pc.ext2);
        S[r] = pk_select(s, pS, keep); G[r] = pk_select(g, pG, keep);
        diag = pS; up = g;
      }
      h = pk_select(bs, h, keep);
      // the node's own column is in the low half for a ref-carrying node, in the high half for a cor-only one
      const uint32_t vS = has_r ? S[R - 1] : pk_swap(S[R - 1]), vG = has_r ? G[R - 1] : pk_swap(G[R - 1]);
      if (!last) { p[R2_BS * 32] = pk_both(vS); p[R2_BG * 32] = pk_both(vG); }
      p[(R2_MOVES + b) * 32] = has_r ? mvl : mvh;
      if (last && (ra & NF_FINAL)) {
        uint32_t v = S[0];
#pragma unroll
        for (int r = 1; r < R; ++r) if (rr == r) v = S[r];
        const int s = (int)((has_r ? v : v >> 16) & 0xffffu) - kBiasP;
        if (s > best) { best = s; best_j = j; }   // ties keep the smaller j (align_lpo_po2.c:410-417)
      }
      ra = ra_n; bs = bs_n; bg = bg_n;
      ra_n = ra_n2; bs_n = bs_n2; bg_n = bg_n2;
    }
  }

  EL_HDN int dp(int nx, int ly, int &best_j) const {
    const int nb = (ly + kBand - 1) / kBand;
    int best = -999999;
    best_j = -1;
    for (int b = 0; b < nb - 1; ++b) band<kBand>(nx, ly, b, false, best, best_j);
    if (ly - (nb - 1) * kBand <= 8) band<8>(nx, ly, nb - 1, true, best, best_j);
    else band<kBand>(nx, ly, nb - 1, true, best, best_j);
    return best;
  }

  // everything up to the traceback; returns the number of MSA columns
  EL_HDN int align_window(const uint16_t *p1, int n1, const uint8_t *unc, int lu, int &s2, AlignBits &al) const {
    fs.pack_codes(sc.tab, unc, lu, Lp->f_unc);
    const int nrings = prepare(p1, n1);
    int bj;
    s2 = dp(n1, lu, bj);
    traceback(n1, lu, bj, al);
    return columns_of(nrings, lu, al.nmatch);
  }
};

}  // namespace elector

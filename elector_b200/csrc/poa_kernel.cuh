// poa_kernel.cuh -- device side of the ELECTOR POA hot path for sm_100a.
//
// One window = (reference, corrected, uncorrected) slices of one read, ~50 bases each
// (SURVEY.md 0.2).  The reference runs, per window (main.c:265-274, buildup_lpo.c:562-589):
//   align_lpo_po(lin(ref), lin(cor)) -> fuse_lpo -> P1
//   align_lpo_po(P1, lin(unc))       -> fuse_lpo -> P2 -> xlate_lpo_to_al (3-row MSA)
// Hundreds of millions of such tiny integer DPs are independent, so the device mapping
// is inter-task: ONE THREAD OWNS ONE WINDOW for the whole pipeline (pack, DP1, traceback,
// fuse, DP2, traceback, fuse, emit), 32 windows of similar size per warp, persistent
// warps pulling 32-window groups from a global counter.  No tensor cores: this is INT32
// DP; the bound is the integer issue rate.
//
// Data placement
//   * DP score columns: shared memory, [row][lane] words -> bank == lane, never a conflict.
//     The DP is swept COLUMN-major over the nodes of the partial order X (x = columns,
//     the growing PO; y = rows, always a linear sequence here).  A node of the 2-sequence
//     PO P1 can only have as predecessors the latest ref-carrying node and the latest
//     cor-carrying node, so at most two "frontier" columns are alive: two buffers per
//     thread, updated in place, replace the reference's (len_y+1) x (len_x+1) matrix.
//   * everything else per window (codes, node records, 2-bit moves, alignment maps, MSA
//     rows) lives in a per-warp scratch area of global memory, interleaved by lane at
//     4-byte granularity so that lock-step lanes produce fully coalesced 128-byte
//     transactions; it is L1/L2 resident and recycled by the persistent warp.
//
// Scoring (align_lpo_po2.c:224-249,384-407 with DOUBLE_GAP_SCORING 0): a cell keeps the
// winning move's score S and gap length g; with a gap table that is flat after the opening
// (pen[0]=open, pen[1..M]=ext; M+1 behaves as 0) only "g != 0" matters.  We store, next
// to S, G = S - pen(g): the value a successor uses for a gap move out of this cell.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace elector {

struct ClassLayout {  // per-thread scratch layout (32-bit words), computed on the host per size class
  int32_t LR, LC, LU;  // caps of the class: max ref / cor / unc length
  int32_t LY;          // rows cap of the shared-memory column buffers = max(LC, LU)
  uint32_t o_ref, o_cor, o_unc;  // packed symbol codes, 4 per word
  uint32_t o_nodeA, o_nodeB;     // node records of the current PO (P0 = lin(ref), then P1)
  uint32_t o_moves;              // 2 bits per DP cell, node-major
  uint32_t o_ord;                // 4 bits per (combined node, row): winning predecessor ordinals
  uint32_t o_x2y, o_y2x;         // alignment maps, one word per entry
  uint32_t o_rows;               // 3 MSA rows, bytes packed 4 per word
  uint32_t o_cols;               // large tier only: the two column buffers, 2 words per entry
  uint32_t row_words;            // words per row in o_rows
  uint32_t ord_wpn;              // words per combined-node slot
  uint32_t total;                // words per thread
};

struct PoaArgs {
  const uint8_t *ref, *cor, *unc;  // raw FASTA letters, concatenated
  const int64_t *ref_off, *cor_off, *unc_off;
  const int32_t *items;  // window ids of this launch (one size class), longest first
  int32_t n_items;
  int32_t match, mismatch, open, ext;
  uint32_t *scratch;
  int32_t *work_counter;
  uint8_t *rows_out;
  unsigned long long *rows_cursor;
  int64_t rows_cap;
  int64_t *row_off;
  int32_t *row_stride, *nring, *score1, *score2;
  int64_t *cells;
  int32_t *error_flag;
  ClassLayout L;
};

// symbol tables: byte -> matrix index (lower-casing + limit_residues + index_symbols),
// index -> output char, and the full substitution table for non-uniform matrices.
struct SymbolTables {
  uint8_t code_lut[256];
  uint8_t sym[32];
  int16_t sub[32 * 32];  // sub[x*32+y] = m->score[x][y]
};

enum : uint32_t {
  NF_REF = 1u << 8,      // node carries a reference letter
  NF_COR = 1u << 9,      // node carries a corrected letter
  NF_INITIAL = 1u << 10, // carries position 0 of some source (align_lpo_po2.c:50-53)
  NF_FINAL = 1u << 11,   // carries the last position of some source (:54-56)
  NF_SAMERING = 1u << 12,// on the same align ring as the previous node
  NF_VIRT = 1u << 13,    // left list starts with the virtual -1 link (:69-75)
  NF_COMB = 1u << 14     // left list has >1 entries: winning ordinals are kept in an o_ord slot
};

#define EL_WARP_FULL 0xffffffffu
#define EL_HD __host__ __device__ __forceinline__
#define EL_HDN __host__ __device__


template <bool GLOBAL_COLS, bool GENERIC_SUB>
struct WindowCtx {
  uint32_t *scr;   // this warp's scratch, indexed [word*32 + lane]
  uint32_t *cols;  // shared: this warp's column buffers (when !GLOBAL_COLS)
  const SymbolTables *tab;
  const ClassLayout *L;
  int lane;
  int match, mismatch, open, ext;
  int colrows;  // rows per column buffer (LY+1 of the class, or ly+1 in the large tier)

  EL_HD uint32_t &sw(uint32_t w) const { return scr[(size_t)w * 32 + lane]; }

  EL_HD int code_at(uint32_t off, int i) const {
    return (sw(off + (i >> 2)) >> ((i & 3) * 8)) & 0xff;
  }

  EL_HD void ld_col(int b, int rr, int &S, int &G) const {
    if (GLOBAL_COLS) {
      uint32_t idx = L->o_cols + 2u * (uint32_t)(b * colrows + rr);
      S = (int)sw(idx);
      G = (int)sw(idx + 1);
    } else {
      uint32_t e = cols[(b * colrows + rr) * 32 + lane];
      S = (int)e >> 16;
      G = (int)(int16_t)(e & 0xffffu);
    }
  }
  EL_HD void st_col(int b, int rr, int S, int G) const {
    if (GLOBAL_COLS) {
      uint32_t idx = L->o_cols + 2u * (uint32_t)(b * colrows + rr);
      sw(idx) = (uint32_t)S;
      sw(idx + 1) = (uint32_t)G;
    } else {
      cols[(b * colrows + rr) * 32 + lane] = ((uint32_t)S << 16) | ((uint32_t)G & 0xffffu);
    }
  }

  // K1: raw letters -> symbol indices, 4 per scratch word
  EL_HDN void pack_codes(const uint8_t *src, int len, uint32_t off) const {
    uint32_t w = 0;
    for (int i = 0; i < len; ++i) {
      w |= (uint32_t)tab->code_lut[src[i]] << ((i & 3) * 8);
      if ((i & 3) == 3) { sw(off + (i >> 2)) = w; w = 0; }
    }
    if (len & 3) sw(off + (len >> 2)) = w;
  }

  // virtual column -1 (align_lpo_po2.c:272-273,290-302): row rr-1
  EL_HD void virt_col(int rr, int &S, int &G) const {
    if (rr == 0) { S = 0; G = -open; }
    else { S = -(open + ext * (rr - 1)); G = S - ext; }
  }

  // ---- DP over the nodes of the current PO (align_lpo_po2.c:269-433), column-major ----
  // nodes: records at o_nodeA (letter | NF_REF/NF_COR/NF_INITIAL/NF_FINAL); this sweep adds
  // NF_VIRT/NF_COMB/slot and writes o_nodeB (the two real predecessors) for the traceback.
  EL_HDN int dp_sweep(int nx, uint32_t o_y, int ly, int &best_j) const {
    const int mwpn = (ly + 15) >> 4;
    int bR = -1, bC = -1, lastR = -1, lastC = -1;
    int best = -999999;
    int nslot = 0;
    best_j = -1;
    for (int j = 0; j < nx; ++j) {
      uint32_t ra = sw(L->o_nodeA + j);
      const int xl = ra & 0xff;
      const bool hasR = ra & NF_REF, hasC = ra & NF_COR;
      int pA = -1, pB = -1, bufA = -1, bufB = -1;
      if (hasR && lastR >= 0) { pA = lastR; bufA = bR; }
      if (hasC && lastC >= 0 && lastC != pA) {
        if (pA < 0) { pA = lastC; bufA = bC; }
        else { pB = lastC; bufB = bC; }
      }
      const bool virt = (ra & NF_INITIAL) && pA >= 0;
      const int nlist = (pA < 0) ? 1 : (int)virt + 1 + (pB >= 0);
      // destination buffer: must not clobber the frontier the node does not carry
      int dst;
      if (hasR && hasC) dst = (bR >= 0) ? bR : ((bC >= 0) ? bC : 0);
      else if (hasR) dst = (bC < 0) ? ((bR >= 0) ? bR : 0) : ((bR >= 0 && bR != bC) ? bR : 1 - bC);
      else dst = (bR < 0) ? ((bC >= 0) ? bC : 0) : ((bC >= 0 && bC != bR) ? bC : 1 - bR);

      int src = bufA;
      if (pA < 0) {  // only the virtual link: materialise column -1
        for (int rr = 0; rr <= ly; ++rr) { int S, G; virt_col(rr, S, G); st_col(dst, rr, S, G); }
        src = dst;
      } else if (nlist > 1) {  // first-strict-max over the left list, per row, with ordinals
        const int slot = nslot++;
        ra |= NF_COMB | ((uint32_t)slot << 16);
        uint32_t ow = 0;
        for (int rr = 0; rr <= ly; ++rr) {
          int bS, bG, oM = 0, oX = 0, k = 0, S, G;
          if (virt) { virt_col(rr, bS, bG); k = 1; ld_col(bufA, rr, S, G); if (S > bS) { bS = S; oM = 1; } if (G > bG) { bG = G; oX = 1; } k = 2; }
          else { ld_col(bufA, rr, bS, bG); k = 1; }
          if (pB >= 0) { ld_col(bufB, rr, S, G); if (S > bS) { bS = S; oM = k; } if (G > bG) { bG = G; oX = k; } }
          st_col(dst, rr, bS, bG);
          ow |= (uint32_t)(oM | (oX << 2)) << ((rr & 7) * 4);
          if ((rr & 7) == 7) { sw(L->o_ord + slot * L->ord_wpn + (rr >> 3)) = ow; ow = 0; }
        }
        if ((ly + 1) & 7) sw(L->o_ord + slot * L->ord_wpn + ((ly + 1) >> 3)) = ow;
        src = dst;
      }
      if (virt) ra |= NF_VIRT;
      sw(L->o_nodeA + j) = ra;
      sw(L->o_nodeB + 2 * j) = (uint32_t)pA;
      sw(L->o_nodeB + 2 * j + 1) = (uint32_t)pB;

      // main column loop
      int pS, pG, S, G, diagS, upG;
      ld_col(src, 0, pS, pG);
      S = pG;           // row -1: gap move out of the predecessor's row -1 (:275-286)
      G = S - ext;
      st_col(dst, 0, S, G);
      diagS = pS; upG = G;
      uint32_t mv = 0, yw = 0;
      const uint32_t mbase = L->o_moves + (uint32_t)j * mwpn;
      for (int r = 0; r < ly; ++r) {
        if ((r & 3) == 0) yw = sw(o_y + (r >> 2));
        const int yc = yw & 0xff; yw >>= 8;
        ld_col(src, r + 1, pS, pG);
        const int sub = GENERIC_SUB ? (int)tab->sub[xl * 32 + yc] : (yc == xl ? match : mismatch);
        const int M = diagS + sub;
        const int gap = pG > upG ? pG : upG;          // ties: Y-gap wins (:392)
        const bool isM = M > gap;              // match must beat both (:384)
        const bool xg = pG > upG;
        S = isM ? M : gap;
        G = S - (isM ? open : ext);
        st_col(dst, r + 1, S, G);
        mv |= ((isM ? 1u : 0u) | (xg ? 2u : 0u)) << ((r & 15) * 2);
        if ((r & 15) == 15) { sw(mbase + (r >> 4)) = mv; mv = 0; }
        diagS = pS; upG = G;
      }
      if (ly & 15) sw(mbase + (ly >> 4)) = mv;
      if ((ra & NF_FINAL) && S > best) { best = S; best_j = j; }  // ties keep the smaller j (:410-417)
      if (hasR) { bR = dst; lastR = j; }
      if (hasC) { bC = dst; lastC = j; }
    }
    return best;
  }

  // ---- traceback (align_lpo_po2.c:108-168) ----
  EL_HDN void traceback(int nx, int ly, int best_j) const {
    const int mwpn = (ly + 15) >> 4;
    for (int j = 0; j < nx; ++j) sw(L->o_x2y + j) = 0xffffffffu;
    for (int r = 0; r < ly; ++r) sw(L->o_y2x + r) = 0xffffffffu;
    int j = best_j, r = ly - 1;
    while (j >= 0 && r >= 0) {
      const uint32_t kind = (sw(L->o_moves + (uint32_t)j * mwpn + (r >> 4)) >> ((r & 15) * 2)) & 3u;
      const uint32_t ra = sw(L->o_nodeA + j);
      if (kind & 1u) { sw(L->o_x2y + j) = (uint32_t)r; sw(L->o_y2x + r) = (uint32_t)j; }
      if ((kind & 1u) || (kind & 2u)) {  // match or X-gap: step to a predecessor of j
        int ord = 0;
        if (ra & NF_COMB) {
          const int rr = (kind & 1u) ? r : r + 1;  // match reads row r-1, X-gap row r
          const uint32_t nib = (sw(L->o_ord + (ra >> 16) * L->ord_wpn + (rr >> 3)) >> ((rr & 7) * 4)) & 15u;
          ord = (kind & 1u) ? (nib & 3u) : (nib >> 2);
        }
        const int pA = (int)sw(L->o_nodeB + 2 * j), pB = (int)sw(L->o_nodeB + 2 * j + 1);
        int nj;
        if (ra & NF_VIRT) nj = (ord == 0) ? -1 : (ord == 1 ? pA : pB);
        else nj = (ord == 0) ? pA : pB;  // pA == -1 when the list is the virtual link alone
        j = nj;
      }
      if ((kind & 1u) || !(kind & 2u)) --r;  // match or Y-gap: step up
    }
  }

  // ---- fuse 1 (lpo.c:413-463,602-656 for two linear sequences): build P1's node records ----
  EL_HDN int fuse1(int lr, int lc) const {
    int n = 0, iy = 0;
    for (int ix = 0; ix < lr; ++ix) {
      const int q = (int)sw(L->o_x2y + ix);
      const int xl = code_at(L->o_ref, ix);
      if (q >= 0)
        while (iy < q) {
          sw(L->o_nodeA + n) = (uint32_t)code_at(L->o_cor, iy) | NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
          ++n; ++iy;
        }
      uint32_t fl = NF_REF | (ix == 0 ? NF_INITIAL : 0u) | (ix == lr - 1 ? NF_FINAL : 0u);
      if (q >= 0 && iy < lc) {
        const int yl = code_at(L->o_cor, iy);
        const uint32_t yf = NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
        if (yl == xl) fl |= yf;  // identical letters share the node
        else { sw(L->o_nodeA + n) = (uint32_t)yl | yf; ++n; fl |= NF_SAMERING; }  // own node just before x, same ring
        ++iy;
      }
      sw(L->o_nodeA + n) = (uint32_t)xl | fl;
      ++n;
    }
    while (iy < lc) {
      sw(L->o_nodeA + n) = (uint32_t)code_at(L->o_cor, iy) | NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
      ++n; ++iy;
    }
    return n;
  }

  // ---- fuse 2 + MSA emit (lpo.c:413-463 with rings, lpo_format.c:346-371) ----
  // Walks the final node order without materialising P2; a column closes whenever the
  // align ring changes.  Returns nring; rows go to o_rows (3 x row_words words).
  EL_HDN int fuse2_emit(int n1, int lu) const {
    const uint8_t *sym = tab->sym;
    int iy = 0, col = -1, prev_key = -1, rs = 0;
    uint32_t c0 = '.', c1 = '.', c2 = '.';
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    const uint32_t r0 = L->o_rows, r1 = L->o_rows + L->row_words, r2 = L->o_rows + 2 * L->row_words;
    auto flush = [&]() {
      if (col >= 0) {
        const int sh = (col & 3) * 8;
        w0 |= c0 << sh; w1 |= c1 << sh; w2 |= c2 << sh;
        if ((col & 3) == 3) { sw(r0 + (col >> 2)) = w0; sw(r1 + (col >> 2)) = w1; sw(r2 + (col >> 2)) = w2; w0 = w1 = w2 = 0; }
      }
    };
    auto node = [&](int key, uint32_t letter, uint32_t srcmask) {
      if (key != prev_key) { flush(); ++col; c0 = c1 = c2 = '.'; prev_key = key; }
      const uint32_t ch = sym[letter];
      if (srcmask & 1u) c0 = ch;
      if (srcmask & 2u) c1 = ch;
      if (srcmask & 4u) c2 = ch;
    };
    for (int ix = 0; ix < n1; ++ix) {
      const uint32_t ra = sw(L->o_nodeA + ix);
      if (!(ra & NF_SAMERING)) rs = ix;
      // scan x's ring from ix on: unaligned y letters go before the first aligned member
      for (int ir = ix;;) {
        const int q = (int)sw(L->o_x2y + ir);
        if (q >= 0) { while (iy < q) { node(n1 + iy, code_at(L->o_unc, iy), 4u); ++iy; } break; }
        ++ir;
        if (ir >= n1 || !(sw(L->o_nodeA + ir) & NF_SAMERING)) break;
      }
      uint32_t mask = ((ra & NF_REF) ? 1u : 0u) | ((ra & NF_COR) ? 2u : 0u);
      if ((int)sw(L->o_x2y + ix) >= 0 && iy < lu) {
        const uint32_t yl = code_at(L->o_unc, iy);
        if (yl == (ra & 0xffu)) mask |= 4u;
        else node(rs, yl, 4u);
        ++iy;
      }
      node(rs, ra & 0xffu, mask);
    }
    while (iy < lu) { node(n1 + iy, code_at(L->o_unc, iy), 4u); ++iy; }
    flush();
    if ((col & 3) != 3) { sw(r0 + (col >> 2)) = w0; sw(r1 + (col >> 2)) = w1; sw(r2 + (col >> 2)) = w2; }
    return col + 1;
  }

  // ---- the whole per-window pipeline (main.c:265-274 + buildup_lpo.c:562-589) ----
  EL_HDN int run_window(const uint8_t *ref, int lr, const uint8_t *cor, int lc, const uint8_t *unc, int lu,
                        int &s1, int &s2, int &n1) const {
    pack_codes(ref, lr, L->o_ref);
    pack_codes(cor, lc, L->o_cor);
    pack_codes(unc, lu, L->o_unc);
    for (int j = 0; j < lr; ++j)  // P0 = lin(ref) (lpo.c:11-32)
      sw(L->o_nodeA + j) = (uint32_t)code_at(L->o_ref, j) | NF_REF | (j == 0 ? NF_INITIAL : 0u) | (j == lr - 1 ? NF_FINAL : 0u);
    int bj;
    s1 = dp_sweep(lr, L->o_cor, lc, bj);
    traceback(lr, lc, bj);
    n1 = fuse1(lr, lc);
    s2 = dp_sweep(n1, L->o_unc, lu, bj);
    traceback(n1, lu, bj);
    return fuse2_emit(n1, lu);
  }
};

// Persistent kernel, one warp per CTA (up to 32 CTAs per SM; shared memory per CTA is what
// bounds residency): each warp repeatedly takes 32 consecutive items of the size-sorted
// work list; lane l owns item base+l.  Shared memory: symbol tables (the 2 KB substitution
// table only for non-uniform matrices) followed by the two column buffers.
template <bool GLOBAL_COLS, bool GENERIC_SUB>
__global__ void __launch_bounds__(32) poa_tpw_kernel(PoaArgs a, const SymbolTables *g_tab) {
  extern __shared__ uint32_t smem[];
  SymbolTables *tab = reinterpret_cast<SymbolTables *>(smem);
  constexpr int kTabWords = (GENERIC_SUB ? sizeof(SymbolTables) : offsetof(SymbolTables, sub)) / 4;
  {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(g_tab);
    for (int i = threadIdx.x; i < kTabWords; i += 32) smem[i] = s[i];
  }
  __syncwarp();
  const int lane = threadIdx.x;
  const size_t warp_slot = blockIdx.x;

  WindowCtx<GLOBAL_COLS, GENERIC_SUB> c;
  c.scr = a.scratch + warp_slot * (size_t)a.L.total * 32;
  c.tab = tab;
  c.L = &a.L;
  c.lane = lane;
  c.match = a.match; c.mismatch = a.mismatch; c.open = a.open; c.ext = a.ext;
  c.colrows = a.L.LY + 1;
  c.cols = smem + kTabWords;

  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(a.work_counter, 32);
    base = __shfl_sync(EL_WARP_FULL, base, 0);
    if (base >= a.n_items) break;
    const bool active = base + lane < a.n_items;
    int nring = 0, w = -1;
    if (active) {
      w = a.items[base + lane];
      const int64_t ro = a.ref_off[w], co = a.cor_off[w], uo = a.unc_off[w];
      const int lr = (int)(a.ref_off[w + 1] - ro), lc = (int)(a.cor_off[w + 1] - co), lu = (int)(a.unc_off[w + 1] - uo);
      if (GLOBAL_COLS) c.colrows = (lc > lu ? lc : lu) + 1;
      int s1, s2, n1;
      nring = c.run_window(a.ref + ro, lr, a.cor + co, lc, a.unc + uo, lu, s1, s2, n1);
      a.nring[w] = nring;
      if (a.score1) a.score1[w] = s1;
      if (a.score2) a.score2[w] = s2;
      if (a.cells) a.cells[w] = (int64_t)lr * lc + (int64_t)n1 * lu;
    }
    __syncwarp();
    // output allocation: warp prefix sum of 3*stride, one atomic per warp
    const int stride = (nring + 3) & ~3;
    const int bytes = 3 * stride;
    int incl = bytes;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(EL_WARP_FULL, incl, d);
      if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(EL_WARP_FULL, incl, 31);
    unsigned long long wbase = 0;
    if (lane == 0) wbase = atomicAdd(a.rows_cursor, (unsigned long long)total);
    wbase = __shfl_sync(EL_WARP_FULL, wbase, 0);
    if (active) {
      const int64_t off = (int64_t)wbase + incl - bytes;
      a.row_off[w] = off;
      a.row_stride[w] = stride;
      if (off + bytes > a.rows_cap) atomicExch(a.error_flag, 1);
      else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.rows_out + off);
        const int sw4 = stride >> 2;
        for (int s = 0; s < 3; ++s)
          for (int k = 0; k < sw4; ++k) dst[s * sw4 + k] = c.sw(a.L.o_rows + s * a.L.row_words + k);
      }
    }
    __syncwarp();
  }
}

}  // namespace elector

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
// DP data placement (the part that decides the speed)
//   * The DP matrix is swept in BANDS of kBand = 8 rows (y = rows, always a linear sequence
//     here; x = columns, the nodes of the growing partial order).  Inside a band a thread
//     keeps the 8 (S, G) cells of the previous column in REGISTERS and updates them in
//     place, fully unrolled: no shared or global memory access per cell.
//   * A node of the 2-sequence PO P1 can only have as predecessors the latest ref-carrying
//     node and the latest cor-carrying node, so at most two "frontier" columns are alive:
//     two register sets A and B (plus which frontier each holds) replace the reference's
//     (len_y+1) x (len_x+1) matrix.  The hot in-place update always runs on set A; the rare
//     nodes that need the other frontier swap / copy / merge the sets first.
//   * Between bands only the band's bottom row travels: one (S, G) pair per node in a
//     per-thread boundary array (global scratch, read once and overwritten once per band).
//   * Moves: 2 bits per cell, 16 bits per (band, node), for the traceback.
//   * Everything per window that is not in registers (codes, node records, moves, boundary
//     row, alignment maps, MSA rows) lives in a per-warp scratch area of global memory,
//     interleaved by lane at 4-byte granularity so that lock-step lanes produce fully
//     coalesced 128-byte transactions; it is L1/L2 resident and recycled by the persistent warp.
//
// Scoring (align_lpo_po2.c:224-249,384-407 with DOUBLE_GAP_SCORING 0): a cell keeps the
// winning move's score S and gap length g; with a gap table that is flat after the opening
// (pen[0]=open, pen[1..M]=ext; M+1 behaves as 0) only "g != 0" matters.  We keep, next
// to S, G = S - pen(g): the value a successor uses for a gap move out of this cell.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace elector {

constexpr int kBand = 8;  // rows per register band

struct ClassLayout {  // per-thread scratch layout (32-bit words), computed on the host per size class
  int32_t LR, LC, LU;  // caps of the class: max ref / cor / unc length
  uint32_t o_ref, o_cor, o_unc;  // packed symbol codes, 4 per word
  uint32_t o_nodeA, o_nodeB;     // node records of the current PO (P0 = lin(ref), then P1)
  uint32_t o_bnd;                // boundary row between bands: (S, G) per node, 2 words
  uint32_t o_moves;              // 16 bits per (band, node): 2 bits per DP cell
  uint32_t o_ord;                // one word per (combined node, band): winning predecessor ordinals
  uint32_t o_x2y, o_y2x;         // alignment maps, one word per entry
  uint32_t o_rows;               // 3 MSA rows, bytes packed 4 per word
  uint32_t row_words;            // words per row in o_rows
  uint32_t total;                // words per thread
};

#define EL_WARP_FULL 0xffffffffu
#define EL_HD __host__ __device__ __forceinline__
#define EL_HDN __host__ __device__

EL_HD uint32_t cdiv_u(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
EL_HD uint32_t max_u(uint32_t a, uint32_t b) { return a > b ? a : b; }

// Scratch layout for windows with ref / cor / unc lengths up to (LR, LC, LU).  Every term is
// monotone in each cap, so a layout for the maxima of a launch bounds the layout of any of
// its 32-window groups (which compute their own, tighter one on the device).
EL_HD void make_layout(ClassLayout &L, int LR, int LC, int LU) {
  L.LR = LR; L.LC = LC; L.LU = LU;
  const uint32_t N1 = (uint32_t)LR + (uint32_t)LC;             // cap of len(P1)
  const uint32_t nb1 = cdiv_u(LC, kBand), nb2 = cdiv_u(LU, kBand);
  uint32_t o = 0;
  L.o_ref = o; o += cdiv_u(LR, 4) + 1;                         // +1: a band reads two code words at once
  L.o_cor = o; o += cdiv_u(LC, 4) + 1;
  L.o_unc = o; o += cdiv_u(LU, 4) + 1;
  L.o_nodeA = o; o += N1;
  L.o_nodeB = o; o += 2 * N1;
  L.o_bnd = o; o += 2 * N1;
  L.o_moves = o; o += cdiv_u(max_u(nb1 * (uint32_t)LR, nb2 * N1), 2);
  L.o_ord = o; o += ((uint32_t)(LR < LC ? LR : LC) + 2) * max_u(nb1, nb2);
  L.o_x2y = o; o += N1;
  L.o_y2x = o; o += max_u(LC, LU);
  L.row_words = cdiv_u(LR + LC + LU, 4);
  L.o_rows = o; o += 3 * L.row_words;
  L.total = o;
}

struct PoaArgs {
  const uint8_t *ref, *cor, *unc;  // raw FASTA letters, concatenated
  const int64_t *ref_off, *cor_off, *unc_off;
  const int32_t *items;  // window ids of this launch (one segment of the size-sorted list), largest first
  int32_t n_items;
  int32_t match, mismatch, open, ext;
  uint32_t *scratch;     // grid x warp_words x 32 words
  uint32_t warp_words;   // scratch words per thread (layout of the segment's maxima)
  int32_t *work_counter;
  uint8_t *rows_out;
  unsigned long long *rows_cursor;
  int64_t rows_cap;
  int64_t *row_off;
  int32_t *row_stride, *nring, *score1, *score2;
  int64_t *cells;
  int32_t *error_flag;
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
  NF_KEEP = 0x1fffu,     // the bits above + the letter: what fuse writes, what survives prepare()
  NF_VIRT = 1u << 13,    // left list starts with the virtual -1 link (:69-75)
  NF_TWO = 1u << 14,     // two real predecessors (latest ref node and latest cor node differ)
  NF_NOPRED = 1u << 15,  // no real predecessor: the left list is the virtual link alone
  NF_PREDC = 1u << 16,   // the single real predecessor is the latest cor-carrying node
  NF_SLOT_SHIFT = 17     // combined nodes (VIRT or TWO): index of their ordinal slot
};

struct ColSet {  // one frontier column inside the current band
  int S[kBand], G[kBand];
  int h;         // S of the row just above the band (row -1 in band 0)
};

template <bool GENERIC_SUB>
struct WindowCtx {
  uint32_t *scr;   // this warp's scratch, indexed [word*32 + lane]
  const SymbolTables *tab;
  const ClassLayout *Lp;  // scratch layout of the current 32-window group (shared memory on the device)
  int lane;
  int match, mismatch, open, ext;

  EL_HD uint32_t &sw(uint32_t w) const { return scr[(size_t)w * 32 + lane]; }

  EL_HD int code_at(uint32_t off, int i) const {
    return (sw(off + (i >> 2)) >> ((i & 3) * 8)) & 0xff;
  }
  EL_HD void st_mv(uint32_t e, uint32_t v) const {
    reinterpret_cast<uint16_t *>(&sw(Lp->o_moves + (e >> 1)))[e & 1] = (uint16_t)v;
  }
  EL_HD uint32_t ld_mv(uint32_t e) const {
    return reinterpret_cast<const uint16_t *>(&sw(Lp->o_moves + (e >> 1)))[e & 1];
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

  // virtual column -1 (align_lpo_po2.c:272-273,290-302) at row `row` (-1 = the corner)
  EL_HD int virt_S(int row) const { return row < 0 ? 0 : -(open + ext * row); }
  EL_HD int virt_G(int row) const { return row < 0 ? -open : -(open + ext * row) - ext; }

  // ---- node preparation (align_lpo_po2.c:46-79 + row -1, :272-286) ----
  // Derives every node's left list from the two frontiers, stores its shape in the node
  // record (NF_VIRT / NF_TWO / NF_NOPRED / NF_PREDC / slot), its real predecessors in
  // o_nodeB (for the traceback) and row -1 of the DP in the boundary array.
  EL_HDN void prepare(int nx) const {
    int lastR = -1, lastC = -1, gR = 0, gC = 0, nslot = 0;
    for (int j = 0; j < nx; ++j) {
      uint32_t ra = sw(Lp->o_nodeA + j) & NF_KEEP;
      const bool hasR = ra & NF_REF, hasC = ra & NF_COR;
      int pA = -1, pB = -1, gA = 0, gB = 0;
      if (hasR && lastR >= 0) { pA = lastR; gA = gR; }
      if (hasC && lastC >= 0 && lastC != pA) {
        if (pA < 0) { pA = lastC; gA = gC; ra |= NF_PREDC; }
        else { pB = lastC; gB = gC; ra |= NF_TWO; }
      }
      const bool virt = (ra & NF_INITIAL) && pA >= 0;
      int bS;  // S(-1, j): first strict maximum of G(-1, p) over the left list
      if (pA < 0) { ra |= NF_NOPRED; bS = -open; }
      else {
        bS = gA;
        if (virt) { ra |= NF_VIRT; bS = -open; if (gA > bS) bS = gA; }
        if (pB >= 0 && gB > bS) bS = gB;
        if (virt || pB >= 0) ra |= (uint32_t)(nslot++) << NF_SLOT_SHIFT;
      }
      const int bG = bS - ext;
      sw(Lp->o_nodeA + j) = ra;
      sw(Lp->o_nodeB + 2 * j) = (uint32_t)pA;
      sw(Lp->o_nodeB + 2 * j + 1) = (uint32_t)pB;
      sw(Lp->o_bnd + 2 * j) = (uint32_t)bS;
      sw(Lp->o_bnd + 2 * j + 1) = (uint32_t)bG;
      if (hasR) { lastR = j; gR = bG; }
      if (hasC) { lastC = j; gC = bG; }
    }
  }

  static EL_HD void swap_sets(ColSet &a, ColSet &b, int &ka, int &kb) {
#pragma unroll
    for (int r = 0; r < kBand; ++r) {
      int t = a.S[r]; a.S[r] = b.S[r]; b.S[r] = t;
      t = a.G[r]; a.G[r] = b.G[r]; b.G[r] = t;
    }
    int t = a.h; a.h = b.h; b.h = t;
    t = ka; ka = kb; kb = t;
  }

  // ---- DP over the nodes of the current PO (align_lpo_po2.c:269-433), band by band ----
  EL_HDN int dp_sweep(int nx, uint32_t o_y, int ly, int &best_j) const {
    prepare(nx);
    const int nb = (ly + kBand - 1) / kBand;
    int best = -999999;
    best_j = -1;
    for (int b = 0; b < nb; ++b) {
      const int r0 = b * kBand;
      int yc[kBand];
      {
        const uint32_t w0 = sw(o_y + (r0 >> 2)), w1 = sw(o_y + (r0 >> 2) + 1);
#pragma unroll
        for (int r = 0; r < kBand; ++r) yc[r] = ((r < 4 ? w0 : w1) >> ((r & 3) * 8)) & 0xff;
      }
      const bool last_band = b == nb - 1;
      ColSet A, B;
#pragma unroll
      for (int r = 0; r < kBand; ++r) A.S[r] = A.G[r] = B.S[r] = B.G[r] = 0;
      A.h = B.h = 0;
      int kindA = 0, kindB = 0;  // which frontiers the sets hold: 1 = ref, 2 = cor, 3 = both
      uint32_t ra = sw(Lp->o_nodeA);
      int upS = (int)sw(Lp->o_bnd), upG = (int)sw(Lp->o_bnd + 1);
      const uint32_t mv_base = (uint32_t)b * (uint32_t)nx;
      for (int j = 0; j < nx; ++j) {
        // prefetch the next node while this one is computed
        const int jn = j + 1 < nx ? j + 1 : j;
        const uint32_t ra_n = sw(Lp->o_nodeA + jn);
        const int upS_n = (int)sw(Lp->o_bnd + 2 * jn), upG_n = (int)sw(Lp->o_bnd + 2 * jn + 1);

        const int m = (ra >> 8) & 3;
        const int xl = ra & 0xff;
        // -- rare: bring the source column into set A, keep the frontier this node leaves behind in B
        if ((ra & (NF_TWO | NF_NOPRED | NF_VIRT | NF_PREDC)) || kindA != m) {
          if (ra & NF_TWO) {
            if (kindA == 2) swap_sets(A, B, kindA, kindB);   // list order: ref predecessor, then cor
          } else if (!(ra & NF_NOPRED)) {
            const int pk = (ra & NF_PREDC) ? 2 : 1;
            if (!(kindA & pk)) swap_sets(A, B, kindA, kindB);
            if (kindA & ~m) { B = A; kindB = kindA & ~m; }
          } else if (kindA & ~m) swap_sets(A, B, kindA, kindB);
          if (ra & NF_NOPRED) {
            A.h = virt_S(r0 - 1);
#pragma unroll
            for (int r = 0; r < kBand; ++r) { A.S[r] = virt_S(r0 + r); A.G[r] = virt_G(r0 + r); }
          } else if (ra & (NF_VIRT | NF_TWO)) {
            // first strict maximum over the left list, per row, S and G separately, with ordinals
            const bool virt = ra & NF_VIRT, two = ra & NF_TWO;
            const uint32_t oA = virt ? 1u : 0u, oB = oA + 1u;
            uint32_t ow = 0;
            {
              int bS = A.h; uint32_t o = oA;
              if (virt) { bS = virt_S(r0 - 1); o = 0; if (A.h > bS) { bS = A.h; o = oA; } }
              if (two && B.h > bS) { bS = B.h; o = oB; }
              A.h = bS; ow |= o;
            }
#pragma unroll
            for (int r = 0; r < kBand; ++r) {
              int bS = A.S[r], bG = A.G[r]; uint32_t oM = oA, oX = oA;
              if (virt) {
                bS = virt_S(r0 + r); bG = virt_G(r0 + r); oM = oX = 0;
                if (A.S[r] > bS) { bS = A.S[r]; oM = oA; }
                if (A.G[r] > bG) { bG = A.G[r]; oX = oA; }
              }
              if (two) {
                if (B.S[r] > bS) { bS = B.S[r]; oM = oB; }
                if (B.G[r] > bG) { bG = B.G[r]; oX = oB; }
              }
              A.S[r] = bS; A.G[r] = bG;
              if (r < kBand - 1) ow |= oM << (2 * (r + 1));
              ow |= oX << (16 + 2 * r);
            }
            sw(Lp->o_ord + (ra >> NF_SLOT_SHIFT) * (uint32_t)nb + b) = ow;
          }
          kindA = m;
          kindB &= ~m;
        }
        // -- hot: in-place update of set A with node j (align_lpo_po2.c:322-407)
        int diag = A.h, up = upG;
        uint32_t mv = 0;
#pragma unroll
        for (int r = 0; r < kBand; ++r) {
          const int pS = A.S[r], pG = A.G[r];
          const int sub = GENERIC_SUB ? (int)tab->sub[xl * 32 + (yc[r] & 31)] : (yc[r] == xl ? match : mismatch);
          const int M = diag + sub;
          const bool xg = pG > up;               // ties: Y-gap beats X-gap (:392)
          const int gap = xg ? pG : up;
          const bool isM = M > gap;              // match must beat both (:384)
          const int s = isM ? M : gap;
          const int g = s - (isM ? open : ext);
          if (isM) mv |= 1u << (2 * r);
          if (xg) mv |= 2u << (2 * r);
          A.S[r] = s; A.G[r] = g;
          diag = pS; up = g;
        }
        A.h = upS;
        sw(Lp->o_bnd + 2 * j) = (uint32_t)A.S[kBand - 1];
        sw(Lp->o_bnd + 2 * j + 1) = (uint32_t)A.G[kBand - 1];
        st_mv(mv_base + j, mv);
        if (last_band && (ra & NF_FINAL)) {
          const int k = (ly - 1) & (kBand - 1);
          int s = A.S[0];
#pragma unroll
          for (int r = 1; r < kBand; ++r) if (k == r) s = A.S[r];
          if (s > best) { best = s; best_j = j; }  // ties keep the smaller j (:410-417)
        }
        ra = ra_n; upS = upS_n; upG = upG_n;
      }
    }
    return best;
  }

  // ---- traceback (align_lpo_po2.c:108-168) ----
  EL_HDN void traceback(int nx, int ly, int best_j) const {
    const int nb = (ly + kBand - 1) / kBand;
    for (int j = 0; j < nx; ++j) sw(Lp->o_x2y + j) = 0xffffffffu;
    for (int r = 0; r < ly; ++r) sw(Lp->o_y2x + r) = 0xffffffffu;
    int j = best_j, r = ly - 1;
    while (j >= 0 && r >= 0) {
      const int b = r >> 3, k = r & 7;
      const uint32_t kind = (ld_mv((uint32_t)b * nx + j) >> (2 * k)) & 3u;
      const uint32_t ra = sw(Lp->o_nodeA + j);
      if (kind & 1u) { sw(Lp->o_x2y + j) = (uint32_t)r; sw(Lp->o_y2x + r) = (uint32_t)j; }
      if (kind) {  // match or X-gap: step to a predecessor of j
        int ord = 0;
        if (ra & (NF_VIRT | NF_TWO)) {
          const uint32_t w = sw(Lp->o_ord + (ra >> NF_SLOT_SHIFT) * (uint32_t)nb + b);
          ord = (int)(((kind & 1u) ? (w >> (2 * k)) : (w >> (16 + 2 * k))) & 3u);
        }
        const int pA = (int)sw(Lp->o_nodeB + 2 * j), pB = (int)sw(Lp->o_nodeB + 2 * j + 1);
        int nj;
        if (ra & NF_VIRT) nj = (ord == 0) ? -1 : (ord == 1 ? pA : pB);
        else nj = (ord == 0) ? pA : pB;  // pA == -1 when the list is the virtual link alone
        j = nj;
      }
      if (kind != 2u) --r;  // match or Y-gap: step up
    }
  }

  // ---- fuse 1 (lpo.c:413-463,602-656 for two linear sequences): build P1's node records ----
  EL_HDN int fuse1(int lr, int lc) const {
    int n = 0, iy = 0;
    for (int ix = 0; ix < lr; ++ix) {
      const int q = (int)sw(Lp->o_x2y + ix);
      const int xl = code_at(Lp->o_ref, ix);
      if (q >= 0)
        while (iy < q) {
          sw(Lp->o_nodeA + n) = (uint32_t)code_at(Lp->o_cor, iy) | NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
          ++n; ++iy;
        }
      uint32_t fl = NF_REF | (ix == 0 ? NF_INITIAL : 0u) | (ix == lr - 1 ? NF_FINAL : 0u);
      if (q >= 0 && iy < lc) {
        const int yl = code_at(Lp->o_cor, iy);
        const uint32_t yf = NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
        if (yl == xl) fl |= yf;  // identical letters share the node
        else { sw(Lp->o_nodeA + n) = (uint32_t)yl | yf; ++n; fl |= NF_SAMERING; }  // own node just before x, same ring
        ++iy;
      }
      sw(Lp->o_nodeA + n) = (uint32_t)xl | fl;
      ++n;
    }
    while (iy < lc) {
      sw(Lp->o_nodeA + n) = (uint32_t)code_at(Lp->o_cor, iy) | NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
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
    const uint32_t r0 = Lp->o_rows, r1 = Lp->o_rows + Lp->row_words, r2 = Lp->o_rows + 2 * Lp->row_words;
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
      const uint32_t ra = sw(Lp->o_nodeA + ix);
      if (!(ra & NF_SAMERING)) rs = ix;
      // scan x's ring from ix on: unaligned y letters go before the first aligned member
      for (int ir = ix;;) {
        const int q = (int)sw(Lp->o_x2y + ir);
        if (q >= 0) { while (iy < q) { node(n1 + iy, code_at(Lp->o_unc, iy), 4u); ++iy; } break; }
        ++ir;
        if (ir >= n1 || !(sw(Lp->o_nodeA + ir) & NF_SAMERING)) break;
      }
      uint32_t mask = ((ra & NF_REF) ? 1u : 0u) | ((ra & NF_COR) ? 2u : 0u);
      if ((int)sw(Lp->o_x2y + ix) >= 0 && iy < lu) {
        const uint32_t yl = code_at(Lp->o_unc, iy);
        if (yl == (ra & 0xffu)) mask |= 4u;
        else node(rs, yl, 4u);
        ++iy;
      }
      node(rs, ra & 0xffu, mask);
    }
    while (iy < lu) { node(n1 + iy, code_at(Lp->o_unc, iy), 4u); ++iy; }
    flush();
    if ((col & 3) != 3) { sw(r0 + (col >> 2)) = w0; sw(r1 + (col >> 2)) = w1; sw(r2 + (col >> 2)) = w2; }
    return col + 1;
  }

  // ---- the whole per-window pipeline (main.c:265-274 + buildup_lpo.c:562-589) ----
  EL_HDN int run_window(const uint8_t *ref, int lr, const uint8_t *cor, int lc, const uint8_t *unc, int lu,
                        int &s1, int &s2, int &n1) const {
    pack_codes(ref, lr, Lp->o_ref);
    pack_codes(cor, lc, Lp->o_cor);
    pack_codes(unc, lu, Lp->o_unc);
    for (int j = 0; j < lr; ++j)  // P0 = lin(ref) (lpo.c:11-32)
      sw(Lp->o_nodeA + j) = (uint32_t)code_at(Lp->o_ref, j) | NF_REF | (j == 0 ? NF_INITIAL : 0u) | (j == lr - 1 ? NF_FINAL : 0u);
    int bj;
    s1 = dp_sweep(lr, Lp->o_cor, lc, bj);
    traceback(lr, lc, bj);
    n1 = fuse1(lr, lc);
    s2 = dp_sweep(n1, Lp->o_unc, lu, bj);
    traceback(n1, lu, bj);
    return fuse2_emit(n1, lu);
  }
};

// Persistent kernel, one warp per CTA (up to 32 CTAs per SM; registers bound residency):
// each warp repeatedly takes 32 consecutive items of the size-sorted work list; lane l owns
// item base+l.  Shared memory holds only the symbol tables (the 2 KB substitution table
// only for non-uniform matrices).
#ifndef EL_MIN_WARPS_PER_SM
#define EL_MIN_WARPS_PER_SM 24  // register cap 80: measured best trade between occupancy and spills
#endif
template <bool GENERIC_SUB>
__global__ void __launch_bounds__(32, EL_MIN_WARPS_PER_SM) poa_tpw_kernel(PoaArgs a, const SymbolTables *g_tab) {
  constexpr int kTabWords = (GENERIC_SUB ? sizeof(SymbolTables) : offsetof(SymbolTables, sub)) / 4;
  __shared__ uint32_t smem[kTabWords];
  __shared__ ClassLayout s_layout;
  SymbolTables *tab = reinterpret_cast<SymbolTables *>(smem);
  {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(g_tab);
    for (int i = threadIdx.x; i < kTabWords; i += 32) smem[i] = s[i];
  }
  __syncwarp();
  const int lane = threadIdx.x;
  const size_t warp_slot = blockIdx.x;

  WindowCtx<GENERIC_SUB> c;
  c.scr = a.scratch + warp_slot * (size_t)a.warp_words * 32;
  c.tab = tab;
  c.Lp = &s_layout;
  c.lane = lane;
  c.match = a.match; c.mismatch = a.mismatch; c.open = a.open; c.ext = a.ext;

  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(a.work_counter, 32);
    base = __shfl_sync(EL_WARP_FULL, base, 0);
    if (base >= a.n_items) break;
    const bool active = base + lane < a.n_items;
    int nring = 0, w = -1, lr = 0, lc = 0, lu = 0;
    int64_t ro = 0, co = 0, uo = 0;
    if (active) {
      w = a.items[base + lane];
      ro = a.ref_off[w]; co = a.cor_off[w]; uo = a.unc_off[w];
      lr = (int)(a.ref_off[w + 1] - ro); lc = (int)(a.cor_off[w + 1] - co); lu = (int)(a.unc_off[w + 1] - uo);
    }
    // the group's own scratch layout: tight, so that its footprint stays in L1/L2
    {
      const int mr = __reduce_max_sync(EL_WARP_FULL, lr), mc = __reduce_max_sync(EL_WARP_FULL, lc), mu = __reduce_max_sync(EL_WARP_FULL, lu);
      __syncwarp();
      if (lane == 0) make_layout(s_layout, mr, mc, mu);
      __syncwarp();
    }
    if (s_layout.total > a.warp_words) { if (lane == 0) atomicExch(a.error_flag, 2); break; }  // cannot happen (monotone layout)
    if (active) {
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
          for (int k = 0; k < sw4; ++k) dst[s * sw4 + k] = c.sw(s_layout.o_rows + s * s_layout.row_words + k);
      }
    }
    __syncwarp();
  }
}

}  // namespace elector

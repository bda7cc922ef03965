// poa_kernel.cuh -- device side of the ELECTOR POA hot path for sm_100a.
//
// One window = (reference, corrected, uncorrected) slices of one read, ~50 bases each
// (SURVEY.md 0.2).  The reference runs, per window (main.c:265-274, buildup_lpo.c:562-589):
//   align_lpo_po(lin(ref), lin(cor)) -> fuse_lpo -> P1
//   align_lpo_po(P1, lin(unc))       -> fuse_lpo -> P2 -> xlate_lpo_to_al (3-row MSA)
// Hundreds of millions of such tiny integer DPs are independent, so the device mapping
// is inter-task: ONE THREAD OWNS ONE WINDOW, 32 windows of the same loop shape per warp,
// persistent warps pulling 32-window groups from a global counter.  No tensor cores: this
// is INT32 DP; the bound is the integer issue rate.
//
// Two kernels, each with its own device-side sort of the windows (bin_kernel.cuh):
//   poa_dp1_kernel : pack ref/cor, DP1 (linear x linear), traceback, fuse 1 -> P1 node list
//                    (16 bits per node, window-major in global memory) + the sort key of phase 2
//   poa_dp2_kernel : pack unc, node preparation, DP2 (P1 x linear), traceback, fuse 2 + MSA emit
// Splitting keeps each kernel's code inside the 32 KB instruction cache (the fused kernel
// stalled on instruction fetch), lets phase 2 be sorted by len(P1) and by WHERE ref and cor
// first differ (lanes of a warp then take the uncommon path of DP2 together instead of one
// after the other), and gives DP1 a smaller register footprint.
//
// DP data placement (the part that decides the speed)
//   * The DP matrix is swept in BANDS of 16 rows (a last band of 8 when at most 8 rows are
//     left); y = rows, always a linear sequence here; x = columns, the nodes of the growing
//     partial order.  Inside a band a thread keeps the (S, G) cells of the previous column
//     in REGISTERS and updates them in place, fully unrolled: no memory access per cell.
//   * DP1 is linear x linear: one register column, predecessor j-1, nothing else.
//   * DP2 runs over P1.  A node of the 2-sequence PO P1 can only have as predecessors the
//     latest ref-carrying node and the latest cor-carrying node, so at most two "frontier"
//     columns are alive: the register set A plus a spare set B in shared memory replace the
//     reference's (len_y+1) x (len_x+1) matrix.  The hot in-place update always runs on A;
//     the few nodes that need the other frontier swap / copy / merge the sets first.
//   * Between bands only the band's bottom row travels: one (S, G) pair per node (global
//     scratch, read once and overwritten once per band).
//   * Moves: 2 bits per cell, one 32-bit word per (band, node), for the traceback.
//   * Everything per window that is not in registers lives in a per-warp scratch area of
//     global memory, interleaved by lane at 4-byte granularity so that lock-step lanes
//     produce fully coalesced 128-byte transactions; per-node data is one record
//     (array of structures) walked with a single running pointer.
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

constexpr int kBand = 16;  // rows per register band (the last band may have 8)

#define EL_WARP_FULL 0xffffffffu
#define EL_HD __host__ __device__ __forceinline__
#define EL_HDN __host__ __device__

EL_HD uint32_t cdiv_u(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
EL_HD uint32_t max_u(uint32_t a, uint32_t b) { return a > b ? a : b; }

// ---- per-thread scratch layouts (32-bit words), one per phase ------------------------------
// Every term is monotone in each cap, so a layout for the maxima of a launch bounds the
// layout of any of its 32-window groups (which compute their own, tighter one on the device).
//
// A layout has a SLOW part (node records: boundary rows and 2-bit moves, written and read once per band by
// the DP with loads issued iterations ahead; always global scratch) and a small FAST part that the serial
// steps of a window walk with dependent accesses (packed letter codes, the alignment bitmaps of the
// traceback): the fast part lives in a per-warp SHARED-MEMORY arena when the group's fits
// (f_total <= PoaArgs::arena_words; ~30-cycle loads instead of 250-600), else in global scratch at o_fast.
// Both are interleaved by lane at 4-byte granularity (word i of a lane at [i * 32]).
enum : uint32_t { R1_BS = 0, R1_BG = 1, R1_MOVES = 2 };                                      // phase-1 node record
enum : uint32_t { R2_NODE = 0, R2_BS = 1, R2_BG = 2, R2_PRED = 3, R2_MOVES = 4 };            // phase-2 node record

struct Layout1 {            // phase 1: windows with ref / cor lengths up to (LR, LC)
  uint32_t o_nodes;         // LR records of rec_words words: boundary S, boundary G, one moves word per band
  uint32_t rec_words;
  uint32_t o_fast;          // the fast part when it lives in global scratch
  uint32_t f_ref, f_cor;    // fast: packed symbol codes, 4 per word
  uint32_t f_xb, f_yb;      // fast: alignment bitmaps (bit j of f_xb: ref letter j is aligned; bit r of f_yb: cor letter r is)
  uint32_t f_total;
  uint32_t total;           // global words per lane (slow + fast)
};
struct Layout2 {            // phase 2: windows with len(P1) / unc length up to (N1, LU)
  uint32_t o_nodes;         // N1 records: node flags|letter, boundary S, boundary G, preds, moves per band
  uint32_t rec_words;
  uint32_t o_ord;           // two words per (combined node, band): winning predecessor ordinals
  uint32_t ord_bands;
  uint32_t o_tmp;           // N1 letter codes (the warp-cooperative kernel rebuilds lin(ref) from them)
  uint32_t o_fast;
  uint32_t f_unc;           // fast: packed symbol codes of unc
  uint32_t f_xb, f_yb;      // fast: alignment bitmaps (node j of P1 / letter r of unc is aligned)
  uint32_t f_nt;            // fast: bitmap, node j has a predecessor list other than [j-1]
  uint32_t f_total;
  uint32_t total;
};

EL_HD void make_layout1(Layout1 &L, int LR, int LC) {
  uint32_t f = 0;
  L.f_ref = f; f += cdiv_u(LR, 4) + 1;
  L.f_cor = f; f += cdiv_u(LC, 4) + 4;                          // +4: a band reads four code words at once
  L.f_xb = f; f += cdiv_u(LR, 32) + 1;
  L.f_yb = f; f += cdiv_u(LC, 32) + 1;
  L.f_total = f;
  uint32_t o = 0;
  L.rec_words = R1_MOVES + cdiv_u(LC, kBand);
  L.o_nodes = o; o += (uint32_t)LR * L.rec_words;
  L.o_fast = o; o += f;
  L.total = o;
}
EL_HD void make_layout2(Layout2 &L, int N1, int LU) {
  const uint32_t nb = cdiv_u(LU, kBand);
  uint32_t f = 0;
  L.f_unc = f; f += cdiv_u(LU, 4) + 4;
  L.f_xb = f; f += cdiv_u(N1, 32) + 1;
  L.f_yb = f; f += cdiv_u(LU, 32) + 1;
  L.f_nt = f; f += cdiv_u(N1, 32) + 1;
  L.f_total = f;
  uint32_t o = 0;
  L.rec_words = R2_MOVES + nb;
  L.o_nodes = o; o += ((uint32_t)N1 + 1) * L.rec_words;         // + 1: the dual kernel's look-ahead loads (node j + 2, j <= N1) stay inside the lane's scratch
  L.ord_bands = nb;
  L.o_ord = o; o += ((uint32_t)N1 / 2 + 2) * nb * 2;            // combined nodes carry ref AND cor: at most N1/2, + 2 initial ones
  L.o_tmp = o; o += cdiv_u(N1, 4) + 1;
  L.o_fast = o; o += f;
  L.total = o;
}

struct PoaArgs {
  const uint8_t *ref, *cor, *unc;  // raw FASTA letters, concatenated
  const int64_t *ref_off, *cor_off, *unc_off;
  // a launch = one segment of a sorted work list; its slice of the list, scratch and work counter come from the segment's
  // device-side plan (tab->plan[seg], bin_kernel.cuh; SegRun below)
  const struct BinTable *tab;
  int32_t seg;
  const int32_t *items_base;   // the sorted work list of the segment's sort
  uint32_t *scratch_base;      // the scratch pool of the segment's phase
  int32_t *ctrl;               // control words of the call (work counters, cursors)
  const long long *rows_cap_dev;   // when set: the end of the segment's row region is read from here instead of rows_cap (two row regions per call)
  int32_t match, mismatch, open, ext;
  uint32_t mis2, nopen2, ext2;   // Scoring::mis2 / nopen2 / ext2
  uint32_t arena_words;  // shared-memory arena words per thread (dynamic shared memory of the launch / 128)
  // phase 1 -> phase 2
  uint16_t *p1_nodes;    // P1 node list of window w at [p1_offset(ref_off[w] - ref_off[0], cor_off[w] - cor_off[0], w)], n1[w] entries
  int64_t ro0, co0;      // ref_off[0], cor_off[0] of this call (offsets may be absolute positions in a larger buffer)
  int32_t *n1;
  int32_t *key2;         // phase-2 sort bin of the window
  int32_t *hist2;        // phase-2 histogram (filled by phase 1)
  int32_t *seg2_max;     // phase-2 segment maxima: [seg*4 + {0: n1, 1: lu}]
  // results
  uint8_t *rows_out;
  unsigned long long *rows_cursor;
  int64_t rows_cap;
  int64_t *row_off;
  int32_t *row_stride, *nring, *score1, *score2;
  int64_t *cells;
  int32_t *error_flag;
  int32_t band_w;        // half-width of the diagonal band of the packed linear kernels (0 = full DP), see BandW
  int32_t band_span;     // band only for groups with len_x + len_y <= this (band_span_limit(maxabs))
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
  NF_KEEP = 0x1fffu,     // the bits above + the letter: what fuse 1 writes (16-bit node list)
  NF_VIRT = 1u << 13,    // left list starts with the virtual -1 link (:69-75)
  NF_TWO = 1u << 14,     // two real predecessors (latest ref node and latest cor node differ)
  NF_NOPRED = 1u << 15,  // no real predecessor: the left list is the virtual link alone
  NF_PREDC = 1u << 16,   // the single real predecessor is the latest cor-carrying node
  NF_SLOT_SHIFT = 17     // combined nodes (VIRT or TWO): index of their ordinal slot
};

// ---- phase-2 sort bins (shared with bin_kernel.cuh) -------------------------------------------
constexpr int kBigTiers = 8;            // windows beyond the "small" limits: one bin per power of two
constexpr int kSmallMax = 256;          // longest sequence of a small window
constexpr int kNbMax = kSmallMax / 8;   // 8-row half-bands of a small window: 1..32
constexpr int kN1q = 128;               // len(P1) / 4 quanta of a small window (len(P1) <= 511)
constexpr int kSpCodes = 98;            // 0 = ref and cor identical; 1 + 3*min(pos/2, 31) + type otherwise
constexpr int kSmallBins2 = kNbMax * kN1q * kSpCodes;
constexpr int kDcls = 4;                 // linear bins: classes of d = len(unc) - len(P1), so that the diagonal bands of a group overlap
constexpr int kLinBins2 = kNbMax * kN1q * kDcls; // small windows whose P1 is linear (ref and cor identical): their own bins and segments
constexpr int kNumBins2 = kBigTiers + kSmallBins2 + kLinBins2;
constexpr int kNumSegs2 = kBigTiers + 8;
constexpr int kFirstLinSeg2 = kBigTiers + 4;

EL_HD int big_tier(int mx) {            // mx > kSmallMax: 1: <= 512, 2: <= 1024, ...
  int t = 1;
  while ((kSmallMax << t) < mx && t < kBigTiers) ++t;
  return t;
}
EL_HD int seg2_of_nb(int nb8) { return kBigTiers + (nb8 > 16 ? 0 : nb8 > 8 ? 1 : nb8 > 4 ? 2 : 3); }
// largest first: big tiers, then the small general bins descending in (half-bands of unc, len(P1)/4, spcode),
// then the small linear bins (spcode 0: DP2 is a linear x linear DP, run by Phase2L) descending in (half-bands, len(P1)/4)
// linear: the window's cor IS its ref (recognised by the size sort, never ran phase 1): Phase2L, sorted and launched
// while phase 1 still runs.  A window that went through phase 1 goes to the general bins whatever its spcode.
EL_HD void bin2_of(int n1, int lu, int spcode, bool linear, int &bin, int &seg) {
  if (lu > kSmallMax || n1 >= 4 * kN1q) {
    const int t = big_tier(n1 > lu ? n1 : lu);
    bin = seg = kBigTiers - t;
  } else {
    const int nb8 = (lu + 7) >> 3;
    if (linear) {
      int dc = (lu - n1 + 6) >> 2;        // d in [-6,-3] [-2,1] [2,5] [6,9]; the tails join the outer classes
      dc = dc < 0 ? 0 : dc > kDcls - 1 ? kDcls - 1 : dc;
      const int lin = ((nb8 - 1) * kN1q + (n1 >> 2)) * kDcls + dc;
      bin = kBigTiers + kSmallBins2 + (kLinBins2 - 1 - lin);
      seg = seg2_of_nb(nb8) + 4;
    } else {
      const int small = ((nb8 - 1) * kN1q + (n1 >> 2)) * kSpCodes + spcode;
      bin = kBigTiers + (kSmallBins2 - 1 - small);
      seg = seg2_of_nb(nb8);
    }
  }
}

// P1 node list of window w: 16-bit entries at this index of PoaArgs::p1_nodes (8-byte aligned, lr+lc+5 entries free)
EL_HD int64_t p1_offset(int64_t ref_off, int64_t cor_off, int64_t w) { return (ref_off + cor_off + 8 * w) & ~(int64_t)3; }

// shifts the sign bit of t into the move word (a negative difference = the move was taken)
EL_HD uint32_t shift_in_sign(uint32_t mv, int t) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l((uint32_t)t, mv, 1);
#else
  return (mv << 1) | ((uint32_t)t >> 31);
#endif
}

// Diagnostic build (-DEL_DP_CLOCKS): SM cycles per phase of the thread-per-window DP2 kernels, summed over warps by lane 0
// (g_dp_clk[kind][phase]; kind 0 = Phase2L, 1 = Phase2D / Phase2; phases: 0 group set-up, 1 pack, 2 node preparation, 3 DP,
// 4 traceback, 5 row allocation + fusion + emit, 6 the rest).  ELECTOR_TRACE prints them after a call (capi.cu).
#ifdef EL_DP_CLOCKS
__device__ unsigned long long g_dp_clk[2][8];
struct PhaseClock {
  mutable long long prev = 0;
  int kind = 0;
  __host__ __device__ void tick(int ph) const {
#ifdef __CUDA_ARCH__
    if ((threadIdx.x & 31) == 0) { const long long t = clock64(); atomicAdd(&g_dp_clk[kind][ph], (unsigned long long)(t - prev)); prev = t; }
#else
    (void)ph;
#endif
  }
};
#define EL_TICK(obj, ph) (obj).pclk.tick(ph)
#else
#define EL_TICK(obj, ph) ((void)0)
#endif

struct Scoring {
  const SymbolTables *tab;
  int match, mismatch, open, ext;
  uint32_t mis2, nopen2, ext2;   // |mismatch|, -open, ext in both 16-bit halves (packed kernels; kernel arguments, so that the
                                 // loops read them from the constant bank instead of rebuilding them)
  EL_HD void set(int m, int mm, int o, int e) {
    match = m; mismatch = mm; open = o; ext = e;
    mis2 = ((uint32_t)(-mm) & 0xffffu) * 0x10001u; nopen2 = ((uint32_t)(-o) & 0xffffu) * 0x10001u; ext2 = ((uint32_t)e & 0xffffu) * 0x10001u;
  }
  // virtual column -1 (align_lpo_po2.c:272-273,290-302) at row `row` (-1 = the corner)
  EL_HD int virt_S(int row) const { return row < 0 ? 0 : -(open + ext * row); }
  EL_HD int virt_G(int row) const { return row < 0 ? -open : -(open + ext * row) - ext; }
};

// ---- diagonal band with an exactness test ----------------------------------------------
// A global path that leaves the band of offsets (row - column) [min(0, d) - w, max(0, d) + w], d = len_y - len_x, has at
// least two gaps with 2w + |d| gap symbols in all, so it scores at most band_bound(sc, w, d) (match scores are <= 0 in the
// packed matrix class).  When the band-restricted DP ends ABOVE that bound every optimal path lies inside the band, the
// cells on optimal paths have their exact values and every comparison the traceback re-reads is decided as in the full
// DP (a predecessor that ties or wins there is on an optimal path itself): score, moves along the path and MSA are the
// full DP's.  Otherwise the window is run again without a band.  Cells outside the band are "minus infinity" = kNegP, a
// value every real score of a small window beats and that cannot leave the 16-bit range by the end of the DP.
constexpr int kBandMaxSpan = 500;           // banding only for groups with len_x + len_y <= this ...
constexpr int kBandMaxScore = 7000;         // ... and maxabs * (len_x + len_y + 4) <= this: real scores stay above kNegP (-8000), whatever the matrix
EL_HD int band_span_limit(int maxabs) { const int s = kBandMaxScore / (maxabs > 0 ? maxabs : 1) - 4; return s < kBandMaxSpan ? s : kBandMaxSpan; }
struct BandW {
  int omin, omax;   // offsets swept (uniform over the warp: a superset of every lane's own band)
  int w;            // half-width the exactness test assumes
  bool on;
};
EL_HD int band_bound(const Scoring &sc, int w, int d) { return -(2 * sc.open + sc.ext * (2 * w + (d < 0 ? -d : d) - 2)); }



// The in-place update of one register column by one node (align_lpo_po2.c:322-407) for R rows:
// S/G hold the predecessor column on entry and the node's column on exit; returns the move bits
// (cell r: bit 2(15-r)+1 = match, bit 2(15-r) = X-gap when not a match).
template <int R, bool GENERIC_SUB>
EL_HD uint32_t update_column(const Scoring &sc, int (&S)[R], int (&G)[R], const uint32_t (&yw)[R / 4], int xl, int diag, int up) {
  uint32_t mv = 0;
  const uint32_t x4 = (uint32_t)xl * 0x01010101u;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pS = S[r], pG = G[r];
    int sub;
    if (GENERIC_SUB) sub = (int)sc.tab->sub[xl * 32 + ((yw[r >> 2] >> ((r & 3) * 8)) & 31)];
    else sub = ((x4 ^ yw[r >> 2]) & (0xffu << ((r & 3) * 8))) ? sc.mismatch : sc.match;
    const int M = diag + sub;
    const int gap = pG > up ? pG : up;     // ties: Y-gap beats X-gap (:392)
    const bool isM = M > gap;              // match must beat both (:384)
    const int s = isM ? M : gap;
    const int g = s - (isM ? sc.open : sc.ext);
    mv = shift_in_sign(mv, gap - M);
    mv = shift_in_sign(mv, up - pG);
    S[r] = s; G[r] = g;
    diag = pS; up = g;
  }
  if (R < kBand) mv <<= 2 * (kBand - R);   // align an 8-row band like the first half of a 16-row one
  return mv;
}

template <int R>
EL_HD int pick_row(const int (&S)[R], int k) {
  int s = S[0];
#pragma unroll
  for (int r = 1; r < R; ++r) if (k == r) s = S[r];
  return s;
}

// scratch access of one lane: word w of the lane at [w*32]
struct LaneScratch {
  uint32_t *base;  // warp scratch + lane
  EL_HD uint32_t &w(uint32_t i) const { return base[(size_t)i * 32]; }
  EL_HD uint32_t *at(uint32_t i) const { return base + (size_t)i * 32; }
  EL_HD int code_at(uint32_t off, int i) const { return (w(off + (i >> 2)) >> ((i & 3) * 8)) & 0xff; }
  // raw letters -> symbol indices, 4 per scratch word.  Reads the letters as aligned 32-bit
  // words (only words that hold at least one letter of the sequence), four loads in flight per step
  // (the letters come from HBM; a one-word look-ahead left this loop waiting on every load).
  EL_HDN void pack_codes(const SymbolTables *tab, const uint8_t *src, int len, uint32_t off) const {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const int mis = (int)(a & 3), sh = mis * 8;
    const int nin = (len + mis + 3) >> 2;   // aligned input words that hold letters
    const int nout = (len + 3) >> 2;
    const uint8_t *lut = tab->code_lut;
    uint32_t cur = wp[0];
#pragma unroll 1
    for (int k = 0; k < nout; k += 4) {
      uint32_t in[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) in[i] = k + 1 + i < nin ? wp[k + 1 + i] : 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t nxt = in[i];
        if (k + i < nout) {
          const uint32_t v = sh ? (cur >> sh) | (nxt << (32 - sh)) : cur;
          uint32_t c = (uint32_t)lut[v & 0xff] | ((uint32_t)lut[(v >> 8) & 0xff] << 8) | ((uint32_t)lut[(v >> 16) & 0xff] << 16) |
                       ((uint32_t)lut[v >> 24] << 24);
          if (k + i == nout - 1 && (len & 3)) c &= 0xffffffffu >> (8 * (4 - (len & 3)));
          w(off + k + i) = c;
        }
        cur = nxt;
      }
    }
  }
};

// brings the line of a scratch word into L1 ahead of a dependent walk (traceback); no register, no wait
EL_HD void prefetch_l1(const void *p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// a store of data this kernel will not read again soon (moves words: re-read once, by the traceback; MSA rows: never):
// evict-first in the L2, so that the boundary rows and node records of the resident groups stay there
EL_HD void st_stream(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}

// ---- what the traceback leaves for the fusion: two bitmaps instead of the reference's x_to_y / y_to_x maps ----
// (align_lpo_po2.c:158-165 fills two integer maps.)  An alignment path is monotone in both sequences: the k-th aligned node
// of x is aligned to the k-th aligned letter of y, so "node j is aligned" / "letter r is aligned" bits carry the same
// information in 1/32 of the space -- small enough for the shared-memory arena.  The walk visits nodes and letters in
// decreasing order and sets the bits in place (the bitmaps are a few words, in shared memory for all but the longest windows).
struct AlignBits {
  LaneScratch st;
  uint32_t ox, oy;
  int nmatch = 0;          // aligned pairs (set by the traceback)
  EL_HD void clear(int nx, int ny) const {
    for (uint32_t k = 0; k < cdiv_u((uint32_t)nx, 32); ++k) st.w(ox + k) = 0;
    for (uint32_t k = 0; k < cdiv_u((uint32_t)ny, 32); ++k) st.w(oy + k) = 0;
  }
  EL_HD void mark(int j, int r) {
    st.w(ox + (uint32_t)(j >> 5)) |= 1u << (j & 31);
    st.w(oy + (uint32_t)(r >> 5)) |= 1u << (r & 31);
    ++nmatch;
  }
  EL_HD bool x_at(int j) const { return (st.w(ox + (uint32_t)(j >> 5)) >> (j & 31)) & 1u; }
};
// where the three MSA rows of a window go: 32-bit words, 4 columns each (the caller's row buffer on the device: the
// fusion writes its output once, in place -- no staging copy)
struct RowSink {
  uint32_t *r0, *r1, *r2;
};

// fuse 1 (lpo.c:413-463,602-656 for two linear sequences): P1's node list, 16 bits per node, to out[]; codes of ref / cor at
// cd.w(o_ref ..) / cd.w(o_cor ..), alignment bitmaps in `al`.  Returns len(P1) and the phase-2 sort code of the window: 0 when
// ref and cor are identical (P1 is linear), else 1 + 3*min(pos/2, 31) + type of the first node that does not
// carry both letters (type 0: ref only, 1: cor only followed by its ref partner = a
// substitution, 2: cor only = an insertion).
// One uniform step per letter of ref or unaligned letter of cor (an unaligned letter of cor goes before the next ALIGNED
// letter of ref, or after the last one): no inner loops, so the lanes of a warp stay in step.
EL_HDN inline int fuse1(const LaneScratch &cd, uint32_t o_ref, uint32_t o_cor, const AlignBits &al, int lr, int lc,
                      uint16_t *out, int &spcode) {
  int n = 0, ix = 0, iy = 0, sp = -1, sptype = 0;
  uint64_t *out4 = reinterpret_cast<uint64_t *>(out);   // 4 nodes per store (the list is 8-byte aligned)
  uint64_t acc = 0;
  auto put = [&](uint32_t v) {
    acc |= (uint64_t)(v & 0xffffu) << (16 * (n & 3));
    if ((n & 3) == 3) { out4[n >> 2] = acc; acc = 0; }
    ++n;
  };
  while (ix < lr || iy < lc) {
    const bool xa = ix < lr && ((al.st.w(al.ox + (uint32_t)(ix >> 5)) >> (ix & 31)) & 1u);
    const bool ya = iy < lc && ((al.st.w(al.oy + (uint32_t)(iy >> 5)) >> (iy & 31)) & 1u);
    const bool yonly = iy < lc && !ya && (ix >= lr || xa);
    const uint32_t xl = (cd.w(o_ref + (uint32_t)(ix >> 2)) >> ((ix & 3) * 8)) & 0xffu;
    const uint32_t yl = (cd.w(o_cor + (uint32_t)(iy >> 2)) >> ((iy & 3) * 8)) & 0xffu;
    const uint32_t yf = NF_COR | (iy == 0 ? NF_INITIAL : 0u) | (iy == lc - 1 ? NF_FINAL : 0u);
    const bool diff = !yonly && xa && yl != xl;                // aligned, different letters: own node just before x, same ring
    if (yonly || diff) {
      if (sp < 0) { sp = n; sptype = yonly ? 2 : 1; }
      put(yl | yf);
    }
    if (!yonly) {
      if (!xa && sp < 0) { sp = n; sptype = 0; }
      put(xl | NF_REF | (ix == 0 ? NF_INITIAL : 0u) | (ix == lr - 1 ? NF_FINAL : 0u) | (xa && !diff ? yf : 0u) | (diff ? NF_SAMERING : 0u));
    }
    ix += !yonly; iy += yonly || xa;
  }
  if (n & 3) out4[n >> 2] = acc;
  spcode = sp < 0 ? 0 : 1 + 3 * ((sp >> 1) < 31 ? (sp >> 1) : 31) + sptype;
  return n;
}

// =============================== phase 1 ========================================================
template <bool GENERIC_SUB>
struct Phase1 {
  typedef Layout1 Layout;
  static constexpr bool kGenericSub = GENERIC_SUB;
  static constexpr bool kBanded = false;
  static EL_HD void make_layout(Layout1 &L, int LR, int LC) { make_layout1(L, LR, LC); }
  LaneScratch scr;    // slow part of the layout (global scratch)
  LaneScratch fs;     // fast part (shared-memory arena, or global scratch at o_fast)
  Scoring sc;
  const Layout1 *Lp;  // layout of the current group (shared memory on the device)

  // DP1: linear x linear (lin(ref) columns, lin(cor) rows), one band
  template <int R>
  EL_HDN int band(int lr, int ly, int b, bool last) const {
    const int r0 = b * kBand;
    uint32_t yw[R / 4];
#pragma unroll
    for (int k = 0; k < R / 4; ++k) yw[k] = fs.w(Lp->f_cor + (r0 >> 2) + k);
    int S[R], G[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { S[r] = sc.virt_S(r0 + r); G[r] = sc.virt_G(r0 + r); }
    int h = sc.virt_S(r0 - 1);
    uint32_t *p = scr.at(Lp->o_nodes);
    const uint32_t step = Lp->rec_words * 32;
    uint32_t xw = 0;
    int upS_n = 0, upG_n = 0;
    if (b > 0) { upS_n = (int)p[R1_BS * 32]; upG_n = (int)p[R1_BG * 32]; }
    for (int j = 0; j < lr; ++j, p += step) {
      if ((j & 3) == 0) xw = fs.w(Lp->f_ref + (j >> 2));
      const int xl = xw & 0xff; xw >>= 8;
      int upS, upG;
      if (b == 0) { upS = -(sc.open + sc.ext * j); upG = upS - sc.ext; }   // row -1 (:275-286)
      else {
        upS = upS_n; upG = upG_n;
        const uint32_t *pn = j + 1 < lr ? p + step : p;
        upS_n = (int)pn[R1_BS * 32]; upG_n = (int)pn[R1_BG * 32];
      }
      const uint32_t mv = update_column<R, GENERIC_SUB>(sc, S, G, yw, xl, h, upG);
      h = upS;
      if (!last) { p[R1_BS * 32] = (uint32_t)S[R - 1]; p[R1_BG * 32] = (uint32_t)G[R - 1]; }
      p[(R1_MOVES + b) * 32] = mv;
    }
    return last ? pick_row<R>(S, ly - 1 - r0) : 0;
  }

  // bands of 16 rows; the last one has 8 when at most 8 rows remain.  P0 = lin(ref)
  // (lpo.c:11-32): only node lr-1 is FINAL, so the best score is the last cell.
  EL_HDN int dp(int lr, int ly) const {
    const int nb = (ly + kBand - 1) / kBand;
    for (int b = 0; b < nb - 1; ++b) band<kBand>(lr, ly, b, false);
    return (ly - (nb - 1) * kBand <= 8) ? band<8>(lr, ly, nb - 1, true) : band<kBand>(lr, ly, nb - 1, true);
  }

  // traceback (align_lpo_po2.c:108-168): marks the aligned pairs in the bitmaps.  The walk reads
  // one moves word per step; the words of the next three columns of the band are loaded ahead.
  EL_HDN void traceback(int lr, int ly, AlignBits &al) const {
    const ptrdiff_t step = (ptrdiff_t)Lp->rec_words * 32;
    al.clear(lr, ly);
    int j = lr - 1, r = ly - 1;
    while (j >= 0 && r >= 0) {
      const int b = r >> 4;
      const uint32_t *pm = scr.at(Lp->o_nodes) + (ptrdiff_t)j * step + (R1_MOVES + b) * 32;
      uint32_t w0 = pm[0], w1 = j >= 1 ? pm[-step] : 0, w2 = j >= 2 ? pm[-2 * step] : 0, w3 = j >= 3 ? pm[-3 * step] : 0;
      for (;;) {
        const uint32_t kind = (w0 >> (2 * (15 - (r & 15)))) & 3u;   // bit 1 match, bit 0 X-gap
        if (kind & 2u) al.mark(j, r);
        if (kind != 1u) --r;
        if (kind) {
          --j; pm -= step;
          w0 = w1; w1 = w2; w2 = w3;
          w3 = j >= 3 ? pm[-3 * step] : 0;
        }
        if (j < 0 || r < 0 || (r >> 4) != b) break;
      }
    }
  }

  EL_HDN int run_window(const uint8_t *ref, int lr, const uint8_t *cor, int lc, uint16_t *p1_out, int &s1, int &spcode, bool &exact) const {
    exact = true;   // no band in the INT32 kernels
    fs.pack_codes(sc.tab, ref, lr, Lp->f_ref);
    fs.pack_codes(sc.tab, cor, lc, Lp->f_cor);
    s1 = dp(lr, lc);
    AlignBits al{fs, Lp->f_xb, Lp->f_yb};
    traceback(lr, lc, al);
    return fuse1(fs, Lp->f_ref, Lp->f_cor, al, lr, lc, p1_out, spcode);
  }
};

// =============================== phase 2 ========================================================
constexpr int kSlotWords = (2 * kBand + 1) * 32;  // one frontier set of a warp in shared memory: S[R], G[R], h, lane-interleaved

// Uncommon nodes of DP2 (first nodes, nodes around a ref/cor difference): the left list is
// not "the column held in set A".  Works on shared-memory copies of both sets, [k*32] per
// lane: sa = set A (holds frontier(s) kindA), sb = set B (kindB); 1 = ref frontier, 2 = cor
// frontier, 3 = both.  On return the node's SOURCE column (first strict maximum over its
// left list, align_lpo_po2.c:334-371, with the winning ordinals in po[0] (match) / po[32]
// (X-gap)) is in A and the frontier the node does not replace is in B.
// Returns kindA | kindB << 2 | 16 if the sets traded places (A is now in sb).
// ordrow0 / ord_or (warp-cooperative DP, poa_coop.cuh): the set covers rows ordrow0 .. ordrow0+R-1 of its 16-row band, and
// a set that does not start the band ORs its ordinals into the words the rows above it wrote one step earlier.
__host__ __device__ __noinline__ inline uint32_t arrange_sets(uint32_t *sa, uint32_t *sb, int R, int r0, uint32_t ra, int kindA,
                                                        int kindB, int open, int ext, uint32_t *po, int ordrow0 = 0, bool ord_or = false) {
  const int m = (ra >> 8) & 3;
  uint32_t flipped = 0;
  auto swap_sets = [&]() {
    uint32_t *t = sa; sa = sb; sb = t;
    const int k = kindA; kindA = kindB; kindB = k;
    flipped ^= 16u;
  };
  auto vS = [&](int row) { return row < 0 ? 0 : -(open + ext * row); };          // virtual column -1
  auto vG = [&](int row) { return row < 0 ? -open : -(open + ext * row) - ext; };
  if (ra & NF_TWO) {
    if (kindA == 2) swap_sets();   // list order: ref predecessor, then cor
  } else if (!(ra & NF_NOPRED)) {
    const int pk = (ra & NF_PREDC) ? 2 : 1;
    if (!(kindA & pk)) swap_sets();
    if (kindA & ~m) {              // the node leaves one of A's frontiers behind: B := A
      for (int k = 0; k <= 2 * R; ++k) sb[k * 32] = sa[k * 32];
      kindB = kindA & ~m;
    }
  } else if (kindA & ~m) swap_sets();
  if (ra & NF_NOPRED) {
    sa[2 * R * 32] = (uint32_t)vS(r0 - 1);
    for (int r = 0; r < R; ++r) { sa[r * 32] = (uint32_t)vS(r0 + r); sa[(R + r) * 32] = (uint32_t)vG(r0 + r); }
  } else if (ra & (NF_VIRT | NF_TWO)) {
    const bool virt = ra & NF_VIRT, two = ra & NF_TWO;
    const uint32_t oA = virt ? 1u : 0u, oB = oA + 1u;
    uint32_t owM = 0, owX = 0;
    {
      int bS = (int)sa[2 * R * 32]; uint32_t o = oA;
      if (virt) { const int a = bS; bS = vS(r0 - 1); o = 0; if (a > bS) { bS = a; o = oA; } }
      if (two) { const int hb = (int)sb[2 * R * 32]; if (hb > bS) { bS = hb; o = oB; } }
      sa[2 * R * 32] = (uint32_t)bS; owM |= o << (2 * ordrow0);
    }
    for (int r = 0; r < R; ++r) {
      const int aS = (int)sa[r * 32], aG = (int)sa[(R + r) * 32];
      int bS = aS, bG = aG; uint32_t oM = oA, oX = oA;
      if (virt) {
        bS = vS(r0 + r); bG = vG(r0 + r); oM = oX = 0;
        if (aS > bS) { bS = aS; oM = oA; }
        if (aG > bG) { bG = aG; oX = oA; }
      }
      if (two) {
        const int sS = (int)sb[r * 32], sG = (int)sb[(R + r) * 32];
        if (sS > bS) { bS = sS; oM = oB; }
        if (sG > bG) { bG = sG; oX = oB; }
      }
      sa[r * 32] = (uint32_t)bS; sa[(R + r) * 32] = (uint32_t)bG;
      const int mr = ordrow0 + r + 1;                  // the row whose match move starts from this cell
      if (mr < kBand) owM |= oM << (2 * mr);           // the band's last row feeds the next band's halo
      owX |= oX << (2 * (ordrow0 + r));
    }
    if (ord_or) { po[0] |= owM; po[32] |= owX; }
    else { po[0] = owM; po[32] = owX; }
  }
  kindA = m;
  kindB &= ~m;
  return (uint32_t)kindA | ((uint32_t)kindB << 2) | flipped;
}

// ---- fuse 2 + MSA emit (lpo.c:413-463 with rings, lpo_format.c:346-371) ----
// Walks the final node order without materialising P2; a column closes whenever the
// align ring changes.  Returns nring; the rows go straight to `out` (4 columns per word).
// The number of columns is known before the walk (columns_of): one per align ring of P1 plus one per
// unaligned letter of unc -- an aligned letter joins its node's ring whether the letters agree or not.
EL_HD int columns_of(int nrings, int lu, int nmatch) { return nrings + lu - nmatch; }

template <class PH>
EL_HDN int fuse_emit_rows(const PH &ph, const AlignBits &al, int n1, int lu, const RowSink &out) {
  const LaneScratch &fs = ph.fs;
  const auto *Lp = ph.Lp;
  const uint8_t *sym = ph.sc.tab->sym;
  int ix = 0, iy = 0, col = -1, prev_key = -1, rs = 0;
  uint32_t c0 = '.', c1 = '.', c2 = '.';
  uint32_t w0 = 0, w1 = 0, w2 = 0;
  auto flush = [&]() {
    if (col >= 0) {
      const int sh = (col & 3) * 8;
      w0 |= c0 << sh; w1 |= c1 << sh; w2 |= c2 << sh;
      if ((col & 3) == 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w1); st_stream(out.r2 + (col >> 2), w2); w0 = w1 = w2 = 0; }
    }
  };
  auto emit = [&](int key, uint32_t letter, uint32_t srcmask) {
    if (key != prev_key) { flush(); ++col; c0 = c1 = c2 = '.'; prev_key = key; }
    const uint32_t ch = sym[letter & 31u];
    if (srcmask & 1u) c0 = ch;
    if (srcmask & 2u) c1 = ch;
    if (srcmask & 4u) c2 = ch;
  };
  const ptrdiff_t step = (ptrdiff_t)Lp->rec_words * 32;
  const uint32_t *pr = ph.node_rec(0);
  // nodes ix, ix+1 in registers, ix+2 in flight.  One uniform step per node of P1 or unaligned letter of unc (which goes
  // before the first aligned member of the next ring that has one, or after the last node): no inner loops.
  uint32_t ra0 = ph.node_flags(pr, 0), ra1 = n1 > 1 ? ph.node_flags(pr + step, 1) : 0, ra2 = n1 > 2 ? ph.node_flags(pr + 2 * step, 2) : 0;
  while (ix < n1 || iy < lu) {
    const bool xa = ix < n1 && al.x_at(ix);
    const bool ya = iy < lu && ((fs.w(al.oy + (uint32_t)(iy >> 5)) >> (iy & 31)) & 1u);
    // is a member of x's ring, from ix on, aligned?  (a ring of P1 has at most two nodes; longer ones are walked all the same)
    bool any = xa;
    if (!any && ix + 1 < n1 && (ra1 & NF_SAMERING)) {
      any = al.x_at(ix + 1);
      for (int ir = ix + 2; !any && ir < n1 && (ph.node_flags(ph.node_rec(ir), ir) & NF_SAMERING); ++ir) any = al.x_at(ir);
    }
    const uint32_t yl = (fs.w(Lp->f_unc + (uint32_t)(iy >> 2)) >> ((iy & 3) * 8)) & 0xffu;
    if (iy < lu && !ya && (ix >= n1 || any)) { emit(n1 + iy, yl, 4u); ++iy; }
    else {
      const uint32_t ra = ra0;
      if (!(ra & NF_SAMERING)) rs = ix;
      uint32_t mask = ((ra & NF_REF) ? 1u : 0u) | ((ra & NF_COR) ? 2u : 0u);
      if (xa) {
        if (yl == (ra & 0xffu)) mask |= 4u;
        else emit(rs, yl, 4u);
        ++iy;
      }
      emit(rs, ra & 0xffu, mask);
      ++ix; pr += step;
      ra0 = ra1; ra1 = ra2;
      ra2 = ix + 2 < n1 ? ph.node_flags(pr + 2 * step, ix + 2) : 0;
    }
  }
  flush();
  if ((col & 3) != 3) { st_stream(out.r0 + (col >> 2), w0); st_stream(out.r1 + (col >> 2), w1); st_stream(out.r2 + (col >> 2), w2); }
  return col + 1;
}

// ---- node preparation (align_lpo_po2.c:46-79 + row -1, :272-286) ----
// Reads P1's 16-bit node list, derives every node's left list from the two frontiers and
// stores: the node with its shape (NF_VIRT / NF_TWO / NF_NOPRED / NF_PREDC / slot), its real
// predecessors (for the traceback) and row -1 of the DP as the first boundary row.  Returns the number of align rings.
template <class PH>
EL_HDN int prepare_nodes(const PH &ph, const uint16_t *nodes, int nx) {
  const auto *Lp = ph.Lp;
  int lastR = -1, lastC = -1, gR = 0, gC = 0, nslot = 0, nrings = 0;
  uint32_t *p = ph.node_rec(0);
  const uint32_t step = Lp->rec_words * 32;
  const int open = ph.sc.open, ext = ph.sc.ext;
  const uint64_t *n4 = reinterpret_cast<const uint64_t *>(nodes);   // 4 nodes per load, one load ahead
  uint64_t quad = n4[0], quad_n = nx > 4 ? n4[1] : 0;
  uint32_t nt = 0;
#pragma unroll 1
  for (int j = 0; j < nx; ++j, p += step) {
    if (j && (j & 3) == 0) { quad = quad_n; quad_n = j + 4 < nx ? n4[(j >> 2) + 1] : 0; }
    uint32_t ra = (uint32_t)(quad >> (16 * (j & 3))) & 0xffffu;
    const bool hasR = ra & NF_REF, hasC = ra & NF_COR;
    if (!(ra & NF_SAMERING)) ++nrings;
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
    p[PH::kRecNode * 32] = ra;
    p[PH::kRecPred * 32] = ((uint32_t)pA & 0xffffu) | ((uint32_t)pB << 16);
    PH::put_row0(p, bS, bG);
    PH::put_shape(p, ra);
    if (pA != j - 1 || (ra & (NF_VIRT | NF_TWO))) nt |= 1u << (j & 31);   // the traceback must look this node up
    if ((j & 31) == 31) { ph.fs.w(Lp->f_nt + (j >> 5)) = nt; nt = 0; }
    if (hasR) { lastR = j; gR = bG; }
    if (hasC) { lastC = j; gC = bG; }
  }
  if (nx & 31) ph.fs.w(Lp->f_nt + (nx >> 5)) = nt;
  return nrings;
}

template <bool GENERIC_SUB>
struct Phase2 {
  typedef Layout2 Layout;
  static constexpr bool kGenericSub = GENERIC_SUB;
  static constexpr bool kLinear = false;
  static constexpr bool kBanded = false;
  static constexpr int kSetWords = kSlotWords;
  static EL_HD void make_layout(Layout2 &L, int N1, int LU) { make_layout2(L, N1, LU); }
  LaneScratch scr;    // slow part of the layout (global scratch)
  LaneScratch fs;     // fast part (shared-memory arena, or global scratch at o_fast)
  uint32_t *bset;     // two frontier-set slots of kSlotWords words, + lane (shared memory on the device)
  Scoring sc;
  const Layout2 *Lp;  // layout of the current group (shared memory on the device)
#ifdef EL_DP_CLOCKS
  PhaseClock pclk;
#endif

  EL_HD uint32_t *rec(uint32_t j) const { return scr.at(Lp->o_nodes + j * Lp->rec_words); }  // field f at [f*32]

  static EL_HD void put_row0(uint32_t *p, int bS, int bG) { p[R2_BS * 32] = (uint32_t)bS; p[R2_BG * 32] = (uint32_t)bG; }
  static EL_HD void put_shape(uint32_t *, uint32_t) {}   // (the dual kernel keeps a shape word per node, poa_dual.cuh)
  EL_HDN int prepare(const uint16_t *nodes, int nx) const { return prepare_nodes(*this, nodes, nx); }

  // ---- DP2: P1 columns x lin(unc) rows, one band (align_lpo_po2.c:269-433) ----
  template <int R>
  EL_HDN void band(int nx, int ly, int b, bool last, int &best, int &best_j) const {
    const int r0 = b * kBand;
    uint32_t yw[R / 4];
#pragma unroll
    for (int k = 0; k < R / 4; ++k) yw[k] = fs.w(Lp->f_unc + (r0 >> 2) + k);
    int S[R], G[R], h = 0;   // set A
#pragma unroll
    for (int r = 0; r < R; ++r) S[r] = G[r] = 0;
    int kindA = 0, kindB = 0;  // which frontiers the sets hold: 1 = ref, 2 = cor, 3 = both
    int bsel = 0;              // which shared-memory slot holds set B
    uint32_t *p = rec(0);
    const uint32_t step = Lp->rec_words * 32;
    uint32_t ra = p[R2_NODE * 32];
    int upS = (int)p[R2_BS * 32], upG = (int)p[R2_BG * 32];
    for (int j = 0; j < nx; ++j, p += step) {
      // prefetch the next node while this one is computed
      const uint32_t *pn = j + 1 < nx ? p + step : p;
      const uint32_t ra_n = pn[R2_NODE * 32];
      const int upS_n = (int)pn[R2_BS * 32], upG_n = (int)pn[R2_BG * 32];

      const int m = (ra >> 8) & 3;
      // -- uncommon: bring the source column into set A, keep the frontier this node leaves behind in B.
      // Done out of line on shared-memory copies of both sets, so that the hot loop stays small.
      if ((ra & (NF_TWO | NF_NOPRED | NF_VIRT | NF_PREDC)) || kindA != m) {
        uint32_t *sa = bset + (1 - bsel) * kSlotWords, *sb = bset + bsel * kSlotWords;
#pragma unroll
        for (int r = 0; r < R; ++r) { sa[r * 32] = (uint32_t)S[r]; sa[(R + r) * 32] = (uint32_t)G[r]; }
        sa[2 * R * 32] = (uint32_t)h;
        uint32_t *po = nullptr;
        if (ra & (NF_VIRT | NF_TWO)) po = scr.at(Lp->o_ord + ((ra >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
        const uint32_t st = arrange_sets(sa, sb, R, r0, ra, kindA, kindB, sc.open, sc.ext, po);
        kindA = st & 3; kindB = (st >> 2) & 3;
        if (st & 16u) { bsel = 1 - bsel; sa = sb; }
#pragma unroll
        for (int r = 0; r < R; ++r) { S[r] = (int)sa[r * 32]; G[r] = (int)sa[(R + r) * 32]; }
        h = (int)sa[2 * R * 32];
      }
      // -- hot: in-place update of set A with node j
      const uint32_t mv = update_column<R, GENERIC_SUB>(sc, S, G, yw, ra & 0xff, h, upG);
      h = upS;
      if (!last) { p[R2_BS * 32] = (uint32_t)S[R - 1]; p[R2_BG * 32] = (uint32_t)G[R - 1]; }
      p[(R2_MOVES + b) * 32] = mv;
      if (last && (ra & NF_FINAL)) {
        const int s = pick_row<R>(S, ly - 1 - r0);
        if (s > best) { best = s; best_j = j; }  // ties keep the smaller j (:410-417)
      }
      ra = ra_n; upS = upS_n; upG = upG_n;
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

  // ---- traceback (align_lpo_po2.c:108-168): marks the aligned pairs in the bitmaps ----
  // One uniform step per cell of the path (load the moves word of (node, band), decode, mark, move; profiles/r2b: a walk
  // with an inner loop per band ran ~4x the instructions at warp level).  Only nodes flagged in the f_nt bitmap need their
  // record looked up; the lines of the cells ahead are requested into L1 as the walk goes.
  EL_HDN void traceback(int nx, int ly, int best_j, AlignBits &al) const {
    const ptrdiff_t step = (ptrdiff_t)Lp->rec_words * 32;
    al.clear(nx, ly);
    int j = best_j, r = ly - 1, nmatch = 0;
    while (j >= 0 && r >= 0) {
      const int b = r >> 4;
      const uint32_t *p = rec((uint32_t)j);
      const uint32_t *pm = p + (R2_MOVES + b) * 32;
      const uint32_t w0 = *pm;
      if (j >= 6) prefetch_l1(pm - 6 * step);
      if (b > 0 && j >= 2) prefetch_l1(pm - 2 * step - 32);
      const uint32_t kind = (w0 >> (2 * (15 - (r & 15)))) & 3u;   // bit 1 match, bit 0 X-gap
      if (kind & 2u) {
        al.st.w(al.ox + (uint32_t)(j >> 5)) |= 1u << (j & 31);
        al.st.w(al.oy + (uint32_t)(r >> 5)) |= 1u << (r & 31);
        ++nmatch;
      }
      if (kind) {  // match or X-gap: step to a predecessor of j
        if ((fs.w(Lp->f_nt + (uint32_t)(j >> 5)) >> (j & 31)) & 1u) {
          const uint32_t ra = p[R2_NODE * 32];
          int ord = 0;
          if (ra & (NF_VIRT | NF_TWO)) {
            const uint32_t *po = scr.at(Lp->o_ord + ((ra >> NF_SLOT_SHIFT) * Lp->ord_bands + (uint32_t)b) * 2);
            ord = (int)((((kind & 2u) ? po[0] : po[32]) >> (2 * (r & 15))) & 3u);
          }
          const uint32_t pr = p[R2_PRED * 32];
          const int pA = (pr & 0xffffu) == 0xffffu ? -1 : (int)(pr & 0xffffu), pB = (pr >> 16) == 0xffffu ? -1 : (int)(pr >> 16);
          if (ra & NF_VIRT) j = (ord == 0) ? -1 : (ord == 1 ? pA : pB);
          else j = (ord == 0) ? pA : pB;  // pA == -1 when the list is the virtual link alone
        } else --j;
      }
      if (kind != 1u) --r;  // match or Y-gap: step up
    }
    al.nmatch = nmatch;
  }

  static constexpr uint32_t kRecNode = R2_NODE, kRecPred = R2_PRED;
  EL_HD uint32_t *node_rec(int j) const { return rec((uint32_t)j); }
  EL_HD uint32_t node_flags(const uint32_t *rec_j, int) const { return rec_j[R2_NODE * 32]; }   // what the fusion reads of node j: letter, NF_REF / NF_COR / NF_SAMERING
  EL_HDN int fuse_emit(const AlignBits &al, int n1, int lu, const RowSink &out) const { return fuse_emit_rows(*this, al, n1, lu, out); }
  EL_HD AlignBits bits() const { return AlignBits{fs, Lp->f_xb, Lp->f_yb}; }

  // everything up to the traceback; returns the number of MSA columns (the rows are emitted once their place is known)
  EL_HDN int align_window(const uint16_t *p1, int n1, const uint8_t *unc, int lu, int &s2, AlignBits &al) const {
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

// ================================ kernels =======================================================
// Persistent kernels, one warp per CTA (up to 32 CTAs per SM; registers bound residency): each
// warp repeatedly takes 32 consecutive items of the sorted work list; lane l owns item base+l.
// Shared memory holds the symbol tables (the 2 KB substitution table only for non-uniform
// matrices), the group's scratch layout and, in phase 2, two frontier-set slots (8.25 KB).
#ifndef EL_MIN_WARPS_PH1
#define EL_MIN_WARPS_PH1 28  // register caps 72 / 96: measured best
#endif
#ifndef EL_MIN_WARPS_PH2
#define EL_MIN_WARPS_PH2 20
#endif

template <bool GENERIC_SUB>
__device__ __forceinline__ const SymbolTables *stage_tables(uint32_t *smem, const SymbolTables *g_tab) {
  constexpr int kTabWords = (GENERIC_SUB ? sizeof(SymbolTables) : offsetof(SymbolTables, sub)) / 4;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(g_tab);
  for (int i = threadIdx.x; i < kTabWords; i += 32) smem[i] = s[i];
  __syncwarp();
  return reinterpret_cast<const SymbolTables *>(smem);
}

#ifdef __CUDACC__
// the band of a group: offsets [min(0, d) - w, max(0, d) + w], d = rows - columns, of every active lane, UNITED over the
// warp.  (Per-lane bands were measured slower than no band at all: lanes at different columns touch different node
// records, and the lane-interleaved scratch is only coalesced when all lanes are at the same column.)
__device__ __forceinline__ void group_band(BandW &bw, bool on, int w, bool active, int d) {
  const int lo = active ? (d < 0 ? d : 0) : 0, hi = active ? (d > 0 ? d : 0) : 0;
  bw.omin = __reduce_min_sync(EL_WARP_FULL, lo) - w;
  bw.omax = __reduce_max_sync(EL_WARP_FULL, hi) + w;
  bw.w = w;
  bw.on = on;
}
// appends the failed lanes' windows to the warp's retry queue; true when 32 of them are ready (then nretry is already
// reduced by 32 and the caller takes queue[nretry + lane])
__device__ __forceinline__ bool queue_retries(int32_t *queue, int &nretry, bool failed, int w) {
  const unsigned fm = __ballot_sync(EL_WARP_FULL, failed);
  if (fm == 0) return false;
  const int lane = threadIdx.x & 31;
  if (failed) queue[nretry + __popc(fm & ((1u << lane) - 1u))] = w;
  nretry += __popc(fm);
  __syncwarp();
  if (nretry < 32) return false;
  nretry -= 32;
  return true;
}

// hist[bin] += 1 for the calling lanes, one atomic per distinct bin among them (neighbouring windows share bins)
__device__ __forceinline__ void warp_hist_add(int32_t *hist, int bin) {
  const unsigned peers = __match_any_sync(__activemask(), bin);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

// What phase 1 leaves for phase 2 (all lanes call it; inactive lanes pass active = false): len(P1), the sort bin, the
// histogram and maxima of the sort of the general segments.
__device__ __forceinline__ void phase1_epilogue(const PoaArgs &a, bool active, int w, int n1, int s1, int spcode, int lr, int lc) {
  (void)lr; (void)lc;
  if (active) {
    const int lu = (int)(a.unc_off[w + 1] - a.unc_off[w]);
    int bin, seg;
    bin2_of(n1, lu, spcode, false, bin, seg);
    a.n1[w] = n1;
    a.key2[w] = bin;
    if (a.score1) a.score1[w] = s1;
    warp_hist_add(a.hist2, bin);
    if (n1 > a.seg2_max[seg * 4]) atomicMax(&a.seg2_max[seg * 4], n1);
    if (lu > a.seg2_max[seg * 4 + 1]) atomicMax(&a.seg2_max[seg * 4 + 1], lu);
  }
}
#endif

// What a CTA reads from its segment's device-side plan (bin_kernel.cuh), kept in SHARED memory: as kernel parameters these
// were free constant-bank operands; as registers they cost the dual-frontier kernel 60 bytes of spills.
struct SegRun {
  const int32_t *items;    // window ids of this launch, largest first
  int32_t n_items;
  uint32_t warp_words;     // scratch words per lane (layout of the segment's maxima)
  uint32_t *scratch;       // max_ctas x warp_words x 32 words
  int32_t *work_counter;
  long long rows_cap;      // end of the segment's row region
};
// false: this CTA has no work
__device__ __forceinline__ bool seg_setup(const PoaArgs &a, SegRun &run);

// the per-warp arena in shared memory (dynamic: the host sizes it so that the kernel's register-bound residency is kept)
extern __shared__ uint32_t s_arena[];
// fast part of the group's layout: the arena when it fits, else global scratch
template <class L>
__device__ __forceinline__ uint32_t *fast_base(const PoaArgs &a, const L &layout, uint32_t *lane_scratch, int lane) {
  return layout.f_total <= a.arena_words ? s_arena + lane : lane_scratch + (size_t)layout.o_fast * 32;
}

// PH = Phase1<GENERIC_SUB> (INT32 cells) or Phase1P (poa_packed.cuh: two 16-bit cells per instruction)
template <class PH, int MIN_WARPS>
__global__ void __launch_bounds__(32, MIN_WARPS) poa_dp1_kernel(PoaArgs a, const SymbolTables *g_tab) {
  constexpr bool GENERIC_SUB = PH::kGenericSub;
  __shared__ uint32_t s_tab[(GENERIC_SUB ? sizeof(SymbolTables) : offsetof(SymbolTables, sub)) / 4];
  __shared__ typename PH::Layout s_layout;
  const int lane = threadIdx.x;
  __shared__ SegRun s_run;
  if (!seg_setup(a, s_run)) return;
  PH c;
  c.scr.base = s_run.scratch + (size_t)blockIdx.x * s_run.warp_words * 32 + lane;
  c.sc.tab = stage_tables<GENERIC_SUB>(s_tab, g_tab);
  c.sc.match = a.match; c.sc.mismatch = a.mismatch; c.sc.open = a.open; c.sc.ext = a.ext;
  c.sc.mis2 = a.mis2; c.sc.nopen2 = a.nopen2; c.sc.ext2 = a.ext2;
  c.Lp = &s_layout;
  __shared__ int32_t s_retry[64];   // windows whose band-restricted DP failed its exactness test: run again without the band
  int nretry = 0;
  // one group: lane l owns window w (active lanes); returns true for a window that has to be run again without the band
  auto process = [&](int w, bool active, bool banded) -> bool {
    int lr = 0, lc = 0;
    int64_t ro = 0, co = 0;
    if (active) {
      ro = a.ref_off[w]; co = a.cor_off[w];
      lr = (int)(a.ref_off[w + 1] - ro); lc = (int)(a.cor_off[w + 1] - co);
    }
    {  // the group's own scratch layout: tight, so that its footprint stays in L1/L2
      const int mr = __reduce_max_sync(EL_WARP_FULL, lr), mc = __reduce_max_sync(EL_WARP_FULL, lc);
      __syncwarp();
      if (lane == 0) PH::make_layout(s_layout, mr, mc);
      __syncwarp();
      // half-width: the base plus 1/16 of the rows (cor differs from ref by ~1 %: scores stay near 0)
      if constexpr (PH::kBanded) group_band(c.bw, banded && mr + mc <= a.band_span, a.band_w + (mc >> 4), active, lc - lr);
    }
    if (s_layout.total > s_run.warp_words) { if (lane == 0) atomicExch(a.error_flag, 2); return false; }  // cannot happen (monotone layout)
    c.fs.base = fast_base(a, s_layout, c.scr.base, lane);
    int s1 = 0, spcode = 0, n1 = 0;
    bool exact = true;
    if (active) n1 = c.run_window(a.ref + ro, lr, a.cor + co, lc, a.p1_nodes + p1_offset(ro - a.ro0, co - a.co0, w), s1, spcode, exact);
    phase1_epilogue(a, active && exact, w, n1, s1, spcode, lr, lc);
    __syncwarp();
    return active && !exact;
  };
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(s_run.work_counter, 32);
    base = __shfl_sync(EL_WARP_FULL, base, 0);
    if (base >= s_run.n_items) break;
    const bool active = base + lane < s_run.n_items;
    const int w = active ? s_run.items[base + lane] : -1;
    const bool failed = process(w, active, a.band_w > 0);
    if constexpr (PH::kBanded) {
      if (queue_retries(s_retry, nretry, failed, w)) { const int w2 = s_retry[nretry + lane]; __syncwarp(); process(w2, true, false); }
    }
  }
  if constexpr (PH::kBanded) {
    if (nretry > 0) { const bool act = lane < nretry; const int w2 = act ? s_retry[lane] : -1; __syncwarp(); process(w2, act, false); }
  }
}

#ifdef __CUDACC__
// Output space for the MSA rows of a group: a warp prefix sum of 3*stride and ONE atomic per warp.  Every lane (active or
// not) calls it; an active lane gets the sink of its window's three rows (nring columns), or false when the caller's row
// buffer is too small (the error flag is then set).
__device__ __forceinline__ bool alloc_window_rows(const PoaArgs &a, long long rows_cap, bool active, int w, int nring, RowSink &out) {
  const int lane = threadIdx.x;
  const int stride = active ? (nring + 3) & ~3 : 0;
  const int bytes = 3 * stride;
  int incl = bytes;
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(EL_WARP_FULL, incl, d);
    if (lane >= d) incl += t;
  }
  const int total = __shfl_sync(EL_WARP_FULL, incl, 31);
  unsigned long long wbase = 0;
  if (lane == 0 && total) wbase = atomicAdd(a.rows_cursor, (unsigned long long)total);
  wbase = __shfl_sync(EL_WARP_FULL, wbase, 0);
  if (!active) return false;
  const int64_t off = (int64_t)wbase + incl - bytes;
  a.row_off[w] = off;
  a.row_stride[w] = stride;
  if (off + bytes > rows_cap) { atomicExch(a.error_flag, 1); return false; }
  out.r0 = reinterpret_cast<uint32_t *>(a.rows_out + off);
  out.r1 = out.r0 + (stride >> 2);
  out.r2 = out.r1 + (stride >> 2);
  return true;
}

#endif
// PH = Phase2<GENERIC_SUB> (INT32 cells), Phase2D (poa_dual.cuh) or Phase2L (poa_packed.cuh)
template <class PH, int MIN_WARPS>
__global__ void __launch_bounds__(32, MIN_WARPS) poa_dp2_kernel(PoaArgs a, const SymbolTables *g_tab) {
  constexpr bool GENERIC_SUB = PH::kGenericSub;
  __shared__ uint32_t s_tab[(GENERIC_SUB ? sizeof(SymbolTables) : offsetof(SymbolTables, sub)) / 4];
  __shared__ typename PH::Layout s_layout;
  __shared__ uint32_t s_bset[2 * PH::kSetWords];
  const int lane = threadIdx.x;
  __shared__ SegRun s_run;
  if (!seg_setup(a, s_run)) return;
  PH c;
  c.scr.base = s_run.scratch + (size_t)blockIdx.x * s_run.warp_words * 32 + lane;
  c.bset = s_bset + lane;
  c.sc.tab = stage_tables<GENERIC_SUB>(s_tab, g_tab);
  c.sc.match = a.match; c.sc.mismatch = a.mismatch; c.sc.open = a.open; c.sc.ext = a.ext;
  c.sc.mis2 = a.mis2; c.sc.nopen2 = a.nopen2; c.sc.ext2 = a.ext2;
  c.Lp = &s_layout;
  __shared__ int32_t s_retry[64];   // windows whose band-restricted DP failed its exactness test: run again without the band
  int nretry = 0;
  // one group: lane l owns window w (active lanes); returns true for a window that has to be run again without the band
  auto process = [&](int w, bool active, bool banded) -> bool {
    int nring = 0, n1 = 0, lu = 0;
    int64_t ro = 0, co = 0, uo = 0;
    if (active) {
      ro = a.ref_off[w]; co = a.cor_off[w]; uo = a.unc_off[w];
      lu = (int)(a.unc_off[w + 1] - uo);
      n1 = a.n1[w];
    }
    {
      const int mn = __reduce_max_sync(EL_WARP_FULL, n1), mu = __reduce_max_sync(EL_WARP_FULL, lu);
      __syncwarp();
      if (lane == 0) PH::make_layout(s_layout, mn, mu);
      __syncwarp();
      // half-width: the base plus 1/8 of the rows (unc differs from ref by ~10 %: the score is about minus the length)
      if constexpr (PH::kBanded) group_band(c.bw, banded && mn + mu <= a.band_span, a.band_w + (mu >> 3), active, lu - n1);
    }
    if (s_layout.total > s_run.warp_words) { if (lane == 0) atomicExch(a.error_flag, 2); return false; }  // cannot happen (monotone layout)
    c.fs.base = fast_base(a, s_layout, c.scr.base, lane);
    bool exact = true;
    AlignBits al = c.bits();
    EL_TICK(c, 0);
    if (active) {
      int s2;
      if constexpr (PH::kLinear) nring = c.align_linear(a.ref + ro, n1, a.unc + uo, lu, s2, exact, al);   // P1 = lin(ref): no node list needed
      else nring = c.align_window(a.p1_nodes + p1_offset(ro - a.ro0, co - a.co0, w), n1, a.unc + uo, lu, s2, al);
      if (exact) {
        a.nring[w] = nring;
        if (a.score2) a.score2[w] = s2;
        if (a.cells) {
          const int64_t lr = a.ref_off[w + 1] - ro, lc = a.cor_off[w + 1] - co;
          a.cells[w] = lr * lc + (int64_t)n1 * lu;
        }
      } else nring = 0;
    }
    __syncwarp();
    RowSink out;
    if (alloc_window_rows(a, s_run.rows_cap, active && exact, w, nring, out)) c.fuse_emit(al, n1, lu, out);
    __syncwarp();
    EL_TICK(c, 5);
    return active && !exact;
  };
#ifdef EL_DP_CLOCKS
  c.pclk.kind = PH::kLinear ? 0 : 1;
  c.pclk.prev = clock64();
#endif
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(s_run.work_counter, 32);
    base = __shfl_sync(EL_WARP_FULL, base, 0);
    if (base >= s_run.n_items) break;
    const bool active = base + lane < s_run.n_items;
    const int w = active ? s_run.items[base + lane] : -1;
    const bool failed = process(w, active, a.band_w > 0);
    if constexpr (PH::kBanded) {
      if (queue_retries(s_retry, nretry, failed, w)) { const int w2 = s_retry[nretry + lane]; __syncwarp(); process(w2, true, false); }
    }
  }
  if constexpr (PH::kBanded) {
    if (nretry > 0) { const bool act = lane < nretry; const int w2 = act ? s_retry[lane] : -1; __syncwarp(); process(w2, act, false); }
  }
}

}  // namespace elector

// tally_kernel.cuh -- device merge (Donatello semantics) + per-read tally (computeStats.py).
//
//   read_totals_kernel / scan_offsets_kernel  : where each read's merged rows start
//   merge_plan_kernel / merge_copy_kernel (Donatello.cpp:13-31,50-84): concatenate a read's window
//                       MSAs, dropping every column whose corrected row is 'n'; plan = one warp per
//                       read (destination of every window), copy = one warp per window
//   tally_read_kernel  one CTA per read: dot bitmasks of the three rows, the sequential scanners of
//                       computeStats.py (:61-98,104-189,472-498: left/right gaps, extension, gap
//                       stretches) run on those bitmasks from dot run to dot run, then the
//                       per-column classification (:291-328,371-440,712-752) under the resulting
//                       mask with a warp-shuffle reduction of the counters
// All are byte streaming kernels (HBM-bound, 3 bytes per MSA column).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/elector_poa.h"

namespace elector {

constexpr int kMaxStretchKeys = 8;
constexpr int T_THRESH = 5, T_THRESH2 = 20;

struct ReadScan {   // output of tally_scan_kernel, input of tally_count_kernel
  int32_t gl, gr;   // gapsLeft / gapsRight (min over reference and uncorrected rows)
  int32_t ext;      // extended bases, -1 if the read is not extended
  int32_t nkeys;    // gap stretches kept by findGapStretches (dict size)
  int32_t key_a[kMaxStretchKeys], key_b[kMaxStretchKeys];
  int32_t overflow; // more than kMaxStretchKeys distinct keys (never seen; reported as an error)
};

// ---- offsets -------------------------------------------------------------------------
// (abort: control word of the call, set when the alignment kernels did not run -- invalid window, scratch pool too small:
// their per-window outputs are then undefined and the merge / tally kernels leave at once)
__global__ void read_totals_kernel(int64_t n_reads, const int64_t *read_first, const int32_t *nring, int64_t *tot, const int32_t *abort) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  if (*abort) { tot[r] = 0; return; }
  int64_t t = 0;
  for (int64_t w = read_first[r]; w < read_first[r + 1]; ++w) t += nring[w];
  tot[r] = (t + 15) & ~(int64_t)15;  // 16-byte aligned slots
}

// single-CTA exclusive scan (n up to a few million: microseconds; avoids a library dependency)
__global__ void __launch_bounds__(1024) scan_offsets_kernel(int64_t n, const int64_t *tot, int64_t *off) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n ? tot[i] : 0;
    int64_t incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int64_t s = warp_sums[lane];
      for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int64_t before = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
    if (i < n) off[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// ---- merge ---------------------------------------------------------------------------
// plan: one warp per read; lanes take the read's windows 32 at a time, count the columns each
// window keeps (corrected row != 'n') and turn them into the window's destination offset.
__global__ void __launch_bounds__(128) merge_plan_kernel(int64_t n_reads, const int64_t *read_first, const uint8_t *rows,
                                                          const int64_t *row_off, const int32_t *row_stride, const int32_t *nring,
                                                          const int64_t *m_off, int64_t *wdst, int32_t *m_len, const int32_t *abort) {
  if (*abort) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_reads) return;
  const int64_t w0 = read_first[r], w1 = read_first[r + 1];
  const int64_t base = m_off[r];
  int64_t done = 0;
  for (int64_t wb = w0; wb < w1; wb += 32) {
    const int64_t w = wb + lane;
    int kept = 0;
    if (w < w1) {
      const int k = nring[w], st = row_stride[w];
      const uint32_t *cor = reinterpret_cast<const uint32_t *>(rows + row_off[w] + st);   // rows are 4-byte aligned, padded with 0
      int dropped = 0;
      for (int i = 0; i < (k + 3) >> 2; ++i) {
        const uint32_t x = cor[i] ^ 0x6e6e6e6eu;                                         // 'n' bytes become 0
        dropped += __popc(~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu));          // exact count of zero bytes
      }
      kept = k - dropped;
    }
    int incl = kept;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (w < w1) wdst[w] = base + done + incl - kept;
    done += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) m_len[r] = (int32_t)done;
}

// copy: a warp takes 32 consecutive windows at a time (their row places, strides, lengths and destinations in one
// coalesced round trip, handed round by shuffles) and copies them four at a time, EIGHT LANES PER WINDOW.  The rows of a
// window are 4-byte aligned and padded, so a lane loads one WORD of each row per pass (4 columns; two passes cover the usual
// 56 columns, all loads of the four windows are issued before the first store).  A window without dropped columns (no 'n' in
// its corrected row: all but the placeholder windows) is a plain copy to a byte-unaligned place: the lanes realign their words
// with a funnel shift against the neighbour's (8-lane shuffles) and store aligned words; the partial words at the two ends are
// handed to lanes 0-3 / 4-7 of the group, one byte each.  Windows that drop columns, or longer than 124 columns, take the
// byte-wise ballot compaction on the whole warp.  (One warp per window with byte accesses made this kernel instruction bound:
// 181 warp instructions per window, ALU pipe 76 %, profiles/r3k_launch_table.csv.)
__device__ __forceinline__ void merge_copy_bytes(const uint8_t *src, int k, int st, int64_t o, int lane, uint8_t *m_ref, uint8_t *m_cor, uint8_t *m_unc) {
  for (int c0 = 0; c0 < k; c0 += 32) {
    const int i = c0 + lane;
    uint8_t c = 'n', a = 0, u = 0;
    if (i < k) { a = src[i]; c = src[st + i]; u = src[2 * st + i]; }
    const unsigned keep = __ballot_sync(0xffffffffu, c != 'n');
    if (c != 'n') {
      const int64_t d = o + __popc(keep & ((1u << lane) - 1u));
      m_ref[d] = a; m_cor[d] = c; m_unc[d] = u;
    }
    o += __popc(keep);
  }
}
__global__ void __launch_bounds__(256) merge_copy_kernel(int64_t n_windows, const uint8_t *__restrict__ rows, const int64_t *__restrict__ row_off,
                                                          const int32_t *__restrict__ row_stride, const int32_t *__restrict__ nring,
                                                          const int64_t *__restrict__ wdst, uint8_t *__restrict__ m_ref,
                                                          uint8_t *__restrict__ m_cor, uint8_t *__restrict__ m_unc, const int32_t *abort) {
  if (*abort) return;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane >> 3, l = lane & 7;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool aligned = ((reinterpret_cast<uintptr_t>(m_ref) | reinterpret_cast<uintptr_t>(m_cor) | reinterpret_cast<uintptr_t>(m_unc)) & 3u) == 0;
  for (int64_t wb = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; wb < n_windows; wb += nwarps * 32) {
    const int64_t wl = wb + lane;
    int k_l = 0, st_l = 0;
    int64_t ro_l = 0, o_l = 0;
    if (wl < n_windows) { k_l = nring[wl]; st_l = row_stride[wl]; ro_l = row_off[wl]; o_l = wdst[wl]; }
    const int nwin = (int)(n_windows - wb < 32 ? n_windows - wb : 32);
    for (int r = 0; 4 * r < nwin; ++r) {
      const int t = 4 * r + g;                                 // my group's window
      int k = __shfl_sync(kFull, k_l, t);
      const int st = __shfl_sync(kFull, st_l, t);
      const int64_t ro = __shfl_sync(kFull, ro_l, t), o = __shfl_sync(kFull, o_l, t);
      if (t >= nwin) k = 0;
      const int a = (int)(o & 3), words = st >> 2;
      bool slow = k > 0 && (!aligned || a + k > 128);
      const int nq = (k > 0 && !slow) ? (a + k + 3) >> 2 : 0;   // destination words of my window (at most 32)
      const int its = (__reduce_max_sync(kFull, (unsigned)nq) + 7) >> 3;
      const uint32_t *s4 = reinterpret_cast<const uint32_t *>(rows + ro);
      uint32_t wa[4], wc[4], wu[4];
      bool hasn = false;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        wa[it] = wc[it] = wu[it] = 0;
        if (it < its) {
          const int q = 8 * it + l;
          if (q < words && nq > 0) { wa[it] = s4[q]; wc[it] = s4[words + q]; wu[it] = s4[2 * words + q]; }
          const uint32_t x = wc[it] ^ 0x6e6e6e6eu;             // 'n' bytes of the corrected row become 0
          hasn |= ((x - 0x01010101u) & ~x & 0x80808080u) != 0;
        }
      }
      if ((__ballot_sync(kFull, hasn) >> (8 * g)) & 0xffu) slow = true;   // my window drops columns
      // destination word q covers the bytes (o - a) + 4q .. + 3 = source bytes 4q - a .. 4q - a + 3
      const int sh = 8 * (4 - a);
      const int64_t d0 = o - a;
      uint32_t ca = 0, cc = 0, cu = 0;                         // source word 8 it - 1
      uint32_t ha = 0, hc = 0, hu = 0, ta = 0, tc = 0, tu = 0; // the first / last destination word, on the lane that made it
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (it < its) {
          uint32_t pa = __shfl_up_sync(kFull, wa[it], 1, 8), pc = __shfl_up_sync(kFull, wc[it], 1, 8), pu = __shfl_up_sync(kFull, wu[it], 1, 8);
          if (l == 0) { pa = ca; pc = cc; pu = cu; }
          ca = __shfl_sync(kFull, wa[it], 7, 8); cc = __shfl_sync(kFull, wc[it], 7, 8); cu = __shfl_sync(kFull, wu[it], 7, 8);
          const uint32_t da = a ? __funnelshift_r(pa, wa[it], sh) : wa[it], dc = a ? __funnelshift_r(pc, wc[it], sh) : wc[it],
                         du = a ? __funnelshift_r(pu, wu[it], sh) : wu[it];
          const int q = 8 * it + l, b0 = 4 * q - a;
          if (!slow && q < nq) {
            if (b0 >= 0 && b0 + 4 <= k) {
              const int64_t d = d0 + 4 * q;
              *reinterpret_cast<uint32_t *>(m_ref + d) = da; *reinterpret_cast<uint32_t *>(m_cor + d) = dc; *reinterpret_cast<uint32_t *>(m_unc + d) = du;
            } else if (q == 0) { ha = da; hc = dc; hu = du; }
            else { ta = da; tc = dc; tu = du; }
          }
        }
      }
      if (its > 0) {   // the partial words at the ends: lanes 0-3 of the group take the bytes of the first word, lanes 4-7 those of the last
        const int qt = nq > 1 ? nq - 1 : 0, lt = qt & 7;
        const uint32_t Ha = __shfl_sync(kFull, ha, 0, 8), Hc = __shfl_sync(kFull, hc, 0, 8), Hu = __shfl_sync(kFull, hu, 0, 8);
        const uint32_t Ta = __shfl_sync(kFull, ta, lt, 8), Tc = __shfl_sync(kFull, tc, lt, 8), Tu = __shfl_sync(kFull, tu, lt, 8);
        const bool tail = l >= 4;
        const int j = l & 3, q = tail ? qt : 0, b = 4 * q - a + j;
        const int bq = 4 * q - a;
        const bool partial = !(bq >= 0 && bq + 4 <= k);        // (a full word went out above)
        if (!slow && nq > 0 && partial && (!tail || qt > 0) && b >= 0 && b < k) {
          const int64_t d = d0 + 4 * q + j;
          m_ref[d] = (uint8_t)((tail ? Ta : Ha) >> (8 * j)); m_cor[d] = (uint8_t)((tail ? Tc : Hc) >> (8 * j)); m_unc[d] = (uint8_t)((tail ? Tu : Hu) >> (8 * j));
        }
      }
      const unsigned slow_groups = __ballot_sync(kFull, slow && l == 0);
      for (int gg = 0; gg < 4; ++gg) {
        if (!((slow_groups >> (8 * gg)) & 1u)) continue;       // uniform
        const int t2 = 4 * r + gg;
        merge_copy_bytes(rows + __shfl_sync(kFull, ro_l, t2), __shfl_sync(kFull, k_l, t2), __shfl_sync(kFull, st_l, t2), __shfl_sync(kFull, o_l, t2), lane, m_ref, m_cor, m_unc);
      }
    }
  }
}

// ---- scanners on dot bitmasks (bit i of word i>>5 = column i is '.'; bits beyond L are 0) ----
struct BitRow {
  const uint32_t *m;
  int L;
  __device__ __forceinline__ bool bit(int i) const { return (m[i >> 5] >> (i & 31)) & 1u; }
  // first column >= i that is a dot, or L
  __device__ int next_set(int i) const {
    while (i < L) {
      const uint32_t w = m[i >> 5] >> (i & 31);
      if (w) return min(L, i + __ffs(w) - 1);
      i = (i | 31) + 1;
    }
    return L;
  }
  // first column p >= i with dots at p and p + 1 (the start of a run of two or more dots when i is not inside a run), or L
  __device__ int next_pair(int i) const {
    const int nw = (L + 31) >> 5;
    while (i < L) {
      const int k = i >> 5;
      const uint32_t w = m[k], nx = k + 1 < nw ? m[k + 1] : 0u;
      const uint32_t p = (w & ((w >> 1) | (nx << 31))) >> (i & 31);
      if (p) return min(L, i + __ffs(p) - 1);
      i = (i | 31) + 1;
    }
    return L;
  }
  // the same for a whole warp that runs the scan redundantly (all lanes pass the same i): 32 words per step
  __device__ int next_pair_warp(int i) const {
    const int nw = (L + 31) >> 5, lane = threadIdx.x & 31;
    if (i >= L) return L;
    const int k0 = i >> 5;
    for (int kb = k0; kb < nw; kb += 32) {
      const int k = kb + lane;
      uint32_t p = 0;
      if (k < nw) {
        const uint32_t w = m[k], nx = k + 1 < nw ? m[k + 1] : 0u;
        p = w & ((w >> 1) | (nx << 31));
        if (k == k0) p &= ~0u << (i & 31);
      }
      const unsigned any = __ballot_sync(0xffffffffu, p != 0);
      if (any) {
        const int src = __ffs(any) - 1;
        const uint32_t pp = __shfl_sync(0xffffffffu, p, src);
        return min(L, (kb + src) * 32 + __ffs(pp) - 1);
      }
    }
    return L;
  }
  // length of the run of columns equal to v that starts at i (i < L), going right
  __device__ int run_fwd(int i, bool v) const {
    const int s = i;
    while (i < L) {
      uint32_t w = m[i >> 5] >> (i & 31);
      if (v) w = ~w;                       // look for the first column that differs
      const int room = 32 - (i & 31);
      const int k = w ? __ffs(w) - 1 : 32;
      if (k < room) return min(L, i + k) - s;
      i += room;
    }
    return L - s;
  }
  // length of the run of columns equal to v that ends at i, going left
  __device__ int run_bwd(int i, bool v) const {
    const int s = i;
    while (i >= 0) {
      uint32_t w = m[i >> 5] << (31 - (i & 31));   // column i at bit 31
      if (v) w = ~w;
      const int room = (i & 31) + 1;
      const int k = w ? __clz(w) : 32;
      if (k < room) return s - (i - k);
      i -= room;
    }
    return s + 1;
  }
  // dots in [a, b)
  __device__ int count(int a, int b) const {
    int n = 0;
    while (a < b) {
      const int hi = min(b, (a | 31) + 1);
      uint32_t w = m[a >> 5] >> (a & 31);
      const int len = hi - a;
      if (len < 32) w &= (1u << len) - 1u;
      n += __popc(w);
      a = hi;
    }
    return n;
  }
};

// nbLeftGaps / nbRightGaps (computeStats.py:61-98): dot runs are consumed whole
__device__ int nb_left_gaps(const BitRow &d) {
  int gaps = 0, nt = 0, total = 0, i = 0;
  while (i < d.L && nt <= T_THRESH) {
    if (d.bit(i)) { const int k = d.run_fwd(i, true); gaps += k; nt = 0; i += k; }
    else { if (gaps >= T_THRESH) total = i; gaps = 0; ++nt; ++i; }
  }
  return total;
}
__device__ int nb_right_gaps(const BitRow &d) {
  int gaps = 0, nt = 0, total = 0, i = d.L - 1;
  while (i >= 0 && nt <= T_THRESH) {
    if (d.bit(i)) { const int k = d.run_bwd(i, true); gaps += k; nt = 0; i -= k; }
    else { if (gaps >= T_THRESH) total = d.L - i; gaps = 0; ++nt; --i; }
  }
  return total;
}

struct StretchState {  // streaming form of findGapStretches' borders / merge / keep passes
  int L;
  int has0, end0;      // dict entry with key 0
  int nkeys;           // dict entries with key != 0 (value is always L-1)
  int key[kMaxStretchKeys];
  int overflow;
  int pend, pa, pb, merge;

  __device__ void emit2(int a, int b) {
    if (a == 0) { if (b - a > T_THRESH2) { has0 = 1; end0 = b; } }
    else if (b == L - 1 && b - a > T_THRESH2) {
      for (int k = 0; k < nkeys; ++k) if (key[k] == a) return;
      if (nkeys < kMaxStretchKeys) key[nkeys++] = a; else overflow = 1;
    }
  }
  __device__ void emit_tmp(int a, int b) {
    if (pend) {
      if (a - pb <= T_THRESH) { emit2(pa, b); merge = 1; }
      else { emit2(pa, pb); merge = 0; }
    }
    pend = 1; pa = a; pb = b;
  }
  __device__ void finalize(int a, int b, bool many) {
    if (many) {
      if (a <= T_THRESH2) emit_tmp(0, b);
      if (L - b <= T_THRESH2) emit_tmp(a, L - 1); else emit_tmp(a, b);
    } else {
      int ea = a <= T_THRESH2 ? 0 : a, eb = b;
      if (L - b <= T_THRESH2) eb = L - 1;
      emit_tmp(ea, eb);
    }
  }
};

// findGapStretches (computeStats.py:104-189) restated over the RUNS of dots of the corrected row.  Run by all 32 lanes of
// one warp with identical state (the search for the next run that counts is the only cooperative step).
// Per column the reference keeps cg (corrected-gap run: 0 on a run's first column unless it is
// column 0, then the run length so far) and cr (reference-gap counter that only advances while the
// previous corrected column was a gap and resets on every reference base).  Only runs that reach
// cg >= 5 can mark columns, a run with cg > 0 closes the current slot when a base follows it, and cr
// at any column follows from the reference-gap run it sits in -- so the scan jumps from run to run.
__device__ void find_gap_stretches(const BitRow &dc, const BitRow &dr, ReadScan &o) {
  const int L = dc.L;
  StretchState s;
  s.L = L; s.has0 = s.end0 = s.nkeys = s.overflow = s.pend = s.pa = s.pb = s.merge = 0;
  int nslots = 0, ca = 0, cb = 0;
  bool have_cur = false, cur_set = false;
  int i = 0;
  while (i < L) {
    // A run of one dot that does not start at column 0 ends with cg == 0: it neither marks columns nor closes a slot.
    // Most runs of the corrected row are such single dots (an inserted base of the uncorrected read), so the scan
    // goes straight to the next run that counts: the one at column 0, or the next run of two or more dots.
    const int st = (i == 0 && dc.bit(0)) ? 0 : dc.next_pair_warp(i);
    if (st >= L) break;
    const int len = dc.run_fwd(st, true), e = st + len - 1;
    const bool counts = st == 0 || len >= 2;            // cg > 0 when the run ends
    if (counts && len >= T_THRESH) {
      // cr after column st (the run's first column never advances cr: its predecessor is a base)
      int cr = 0;
      if (dr.bit(st)) {
        const int sr = st - dr.run_bwd(st, true) + 1;    // start of the reference-gap run holding st
        const int k = dc.count(max(sr, 1) - 1, st);
        cr = sr == 0 ? 1 + k : (k == 0 ? 0 : k + 1);
      }
      int first = -1, last = -1;
      int p = st + 1;
      while (p <= e) {
        const int wend = min(e, p | 31);
        uint32_t rb = dr.m[p >> 5] >> (p & 31);
        const int n = wend - p + 1;
        if (n < 32) rb &= (1u << n) - 1u;
        if (rb == 0) {                                    // reference bases only: cr = 0 on all of them
          cr = 0;
          const int lo = max(p, st + T_THRESH - 1);
          if (lo <= wend) { if (first < 0) first = lo; last = wend; }
        } else {
          for (int q = p; q <= wend; ++q) {
            if (dr.bit(q)) cr = cr > 0 ? cr + 1 : 2; else cr = 0;
            if (q >= st + T_THRESH - 1 && cr < T_THRESH2) { if (first < 0) first = q; last = q; }
          }
        }
        p = wend + 1;
      }
      if (first >= 0) {
        if (nslots == 0) { have_cur = true; cur_set = true; nslots = 1; ca = first - T_THRESH + 1; cb = last; }
        else { if (!cur_set) { cur_set = true; ca = first - T_THRESH + 1; } cb = last; }
      }
    }
    if (counts && e + 1 < L) {                            // a base follows: the run closes the current slot
      if (have_cur && cur_set) s.finalize(ca, cb, true);
      have_cur = true; cur_set = false; ++nslots;
    }
    i = e + 1;
  }
  if (have_cur && cur_set) s.finalize(ca, cb, nslots > 1);
  if (s.pend && !s.merge) s.emit2(s.pa, s.pb);
  if (s.has0) { o.key_a[o.nkeys] = 0; o.key_b[o.nkeys] = s.end0; ++o.nkeys; }
  for (int k = 0; k < s.nkeys && o.nkeys < kMaxStretchKeys; ++k) { o.key_a[o.nkeys] = s.key[k]; o.key_b[o.nkeys] = L - 1; ++o.nkeys; }
  o.overflow = s.overflow || (s.has0 + s.nkeys > kMaxStretchKeys);
}

// ---- the per-read tally: one CTA per read ------------------------------------------------------
//   A  all threads: one pass over the three rows -> dot bitmasks (3 bits per column).  The masks of a
//      read of up to kSmemCols columns live in shared memory; longer reads use the global planes.
//   B  thread 0: the sequential scanners of computeStats.py on the bitmasks (run to run, not column
//      to column): left / right gaps, extension, gap stretches -> the column mask
//   C  all threads: second pass, per-column classification under the mask (computeStats.py:291-328,
//      371-440, 712-752), warp-shuffle + shared-memory reduction of the counters
// Rows whose start is 16-byte aligned (the merge kernels always produce such rows) are read as
// 16-byte vectors, 16 columns per thread and load; other rows take the byte path.
constexpr int kNAcc = 19;  // accumulators reduced per read (see below)
constexpr int kSmemCols = 32768;
constexpr int kSmemMaskWords = kSmemCols / 32;

// 4 bytes -> 4 bits: bit k set when byte k of x equals the byte replicated in pat
__device__ __forceinline__ uint32_t eq_nibble(uint32_t x, uint32_t pat) {
  const uint32_t z = x ^ pat;
  const uint32_t m = ~(((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z | 0x7f7f7f7fu);   // 0x80 in every zero byte of z
  return (((m >> 7) * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4 &v, uint32_t pat) {
  return eq_nibble(v.x, pat) | (eq_nibble(v.y, pat) << 4) | (eq_nibble(v.z, pat) << 8) | (eq_nibble(v.w, pat) << 12);
}

// 4 bytes -> 4 bits: bit k set when byte k of z is zero
__device__ __forceinline__ uint32_t zero_nibble(uint32_t z) {
  const uint32_t m = ~(((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z | 0x7f7f7f7fu);
  return (((m >> 7) * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ uint32_t eq_pair16(const uint4 &a, const uint4 &b) {   // bit k: byte k of a equals byte k of b
  return zero_nibble(a.x ^ b.x) | (zero_nibble(a.y ^ b.y) << 4) | (zero_nibble(a.z ^ b.z) << 8) | (zero_nibble(a.w ^ b.w) << 12);
}
__device__ __forceinline__ uint32_t gc_mask16(const uint4 &v) {                    // bit k: byte k is one of g c G C
  const uint4 f = make_uint4(v.x | 0x20202020u, v.y | 0x20202020u, v.z | 0x20202020u, v.w | 0x20202020u);
  return eq_mask16(f, 0x67676767u) | eq_mask16(f, 0x63636363u);
}
__device__ __forceinline__ uint32_t range_mask16(int lo, int hi) {                 // bits lo..hi (clamped to 0..15), empty when hi < lo
  lo = max(lo, 0); hi = min(hi, 15);
  return hi < lo ? 0u : ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}

struct ColumnMask {   // which columns the classification skips (gapsAndExtensions + gap stretches)
  int lmask, rmask, nkeys;
  int key_a[kMaxStretchKeys], key_b[kMaxStretchKeys];
};

// classification of one column (computeStats.py:291-328,371-440,712-752)
__device__ __forceinline__ void classify_column(int i, uint32_t r_, uint32_t c_, uint32_t u_, const ColumnMask &cm, int (&acc)[kNAcc]) {
  acc[13] += (r_ == 'g' || r_ == 'c' || r_ == 'G' || r_ == 'C');
  acc[14] += (c_ == 'g' || c_ == 'c' || c_ == 'G' || c_ == 'C');
  acc[15] += r_ == '.'; acc[16] += c_ == '.'; acc[17] += u_ == '.';
  bool ok = i >= cm.lmask && i <= cm.rmask;
  for (int k = 0; k < cm.nkeys; ++k) {
    const bool in = i >= cm.key_a[k] && i <= cm.key_b[k];
    if (in) { ok = false; acc[18] += (r_ == '.'); }           // dots of the reference row inside stretches
  }
  if (!ok) return;
  if (c_ != r_) { if (r_ == '.') ++acc[7]; else if (c_ != '.') ++acc[9]; else ++acc[8]; }
  if (u_ != r_) { if (r_ == '.') ++acc[10]; else if (u_ != '.') ++acc[12]; else ++acc[11]; }
  if (r_ == u_) {
    if (u_ != c_) { ++acc[1]; ++acc[4]; } else { ++acc[0]; ++acc[3]; }
    ++acc[5];
  } else {
    if (r_ == c_) { ++acc[0]; ++acc[3]; }
    else { if (u_ == c_) { ++acc[2]; ++acc[1]; } ++acc[4]; }
    ++acc[6];
  }
}

#ifdef ELECTOR_TALLY_TIMING
__device__ unsigned long long g_tally_clk[4];
#define TT(k) if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_tally_clk[k], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
#else
#define TT(k)
#endif
__global__ void __launch_bounds__(128) tally_read_kernel(int64_t n_reads, const uint8_t *R, const uint8_t *C, const uint8_t *U,
                                                          const int64_t *off, const int32_t *len, uint32_t *bits, int64_t plane_words,
                                                          int64_t *counters, int32_t *stretches, int32_t *overflow_flag, const int32_t *abort) {
  const int64_t r = blockIdx.x;
  if (r >= n_reads || *abort) return;
  const int L = len ? len[r] : (int)(off[r + 1] - off[r]);
  const uint8_t *rr = R + off[r], *cc = C + off[r], *uu = U + off[r];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool assessed = L > 10;
  const bool vec = (off[r] & 15) == 0;
#ifdef ELECTOR_TALLY_TIMING
  long long t_prev = clock64();
#endif
  __shared__ ReadScan sc;
  __shared__ int sm[4][kNAcc];
  __shared__ uint32_t s_bits[3 * kSmemMaskWords];
  uint32_t *br, *bc, *bu;
  if (L <= kSmemCols) { br = s_bits; bc = s_bits + kSmemMaskWords; bu = s_bits + 2 * kSmemMaskWords; }
  else { br = bits + (off[r] >> 5) + r; bc = br + plane_words; bu = bc + plane_words; }
  if (assessed) {                                        // A
    if (vec) {
      const uint4 *r4 = reinterpret_cast<const uint4 *>(rr), *c4 = reinterpret_cast<const uint4 *>(cc), *u4 = reinterpret_cast<const uint4 *>(uu);
      const int nq = (L + 15) >> 4, nq2 = (nq + 1) & ~1;   // 16-column groups; lanes work in pairs (one mask word per pair)
      for (int q = threadIdx.x; q < ((nq2 + 127) & ~127); q += 128) {
        uint32_t mr = 0, mc = 0, mu = 0;
        if (q < nq) {
          const int left = L - q * 16;
          const uint32_t valid = left >= 16 ? 0xffffu : (1u << left) - 1u;
          mr = eq_mask16(r4[q], 0x2e2e2e2eu) & valid;
          mc = eq_mask16(c4[q], 0x2e2e2e2eu) & valid;
          mu = eq_mask16(u4[q], 0x2e2e2e2eu) & valid;
        }
        const uint32_t pr = __shfl_down_sync(0xffffffffu, mr, 1), pc = __shfl_down_sync(0xffffffffu, mc, 1), pu = __shfl_down_sync(0xffffffffu, mu, 1);
        if (!(q & 1) && q < nq2) { br[q >> 1] = mr | (pr << 16); bc[q >> 1] = mc | (pc << 16); bu[q >> 1] = mu | (pu << 16); }
      }
    } else {
      for (int base = wid * 32; base < L; base += 128) {
        const int i = base + lane;
        const bool in = i < L;
        const unsigned mr = __ballot_sync(0xffffffffu, in && rr[i] == '.');
        const unsigned mc = __ballot_sync(0xffffffffu, in && cc[i] == '.');
        const unsigned mu = __ballot_sync(0xffffffffu, in && uu[i] == '.');
        if (lane == 0) { br[base >> 5] = mr; bc[base >> 5] = mc; bu[base >> 5] = mu; }
      }
    }
  }
  __syncthreads();
  TT(0)
  if (threadIdx.x < 32) {                                // B (warp 0, every lane with the same state)
    ReadScan o;
    o.gl = o.gr = 0; o.ext = -1; o.nkeys = 0; o.overflow = 0;
    for (int k = 0; k < kMaxStretchKeys; ++k) o.key_a[k] = o.key_b[k] = 0;
    if (assessed) {
      const BitRow dr{br, L}, dc{bc, L}, du{bu, L};
      o.gl = min(nb_left_gaps(dr), nb_left_gaps(du));
      o.gr = min(nb_right_gaps(dr), nb_right_gaps(du));
      int ext = -1;
      if (o.gl >= T_THRESH2) ext = (ext < 0 ? 0 : ext) + o.gl - dc.count(0, o.gl);
      if (o.gr >= T_THRESH2) ext = (ext < 0 ? 0 : ext) + o.gr - dc.count(L - o.gr + 1, L);
      o.ext = ext;
      find_gap_stretches(dc, dr, o);
      if (o.overflow && threadIdx.x == 0) atomicExch(overflow_flag, 1);
    }
    if (threadIdx.x == 0) sc = o;
  }
  __syncthreads();
  TT(1)
  int acc[kNAcc];                                        // C
#pragma unroll
  for (int k = 0; k < kNAcc; ++k) acc[k] = 0;
  ColumnMask cm;
  cm.lmask = sc.gl >= T_THRESH ? sc.gl : 0;                        // columns [0, gl) masked
  cm.rmask = sc.gr >= T_THRESH ? L - sc.gr : L - 1;                // columns (L-gr, L-1] masked
  cm.nkeys = sc.nkeys;
#pragma unroll
  for (int k = 0; k < kMaxStretchKeys; ++k) { cm.key_a[k] = sc.key_a[k]; cm.key_b[k] = sc.key_b[k]; }
  if (assessed) {
    if (vec) {
      const uint4 *r4 = reinterpret_cast<const uint4 *>(rr), *c4 = reinterpret_cast<const uint4 *>(cc), *u4 = reinterpret_cast<const uint4 *>(uu);
      const int nq = (L + 15) >> 4;
      // 16 columns at a time as bitmasks: byte equalities (SWAR), dots, the column mask; every counter is a popcount
      for (int q = threadIdx.x; q < nq; q += 128) {
        const uint4 vr = r4[q], vc = c4[q], vu = u4[q];
        const int i0 = q * 16;
        const uint32_t V = range_mask16(0, L - 1 - i0);
        const uint32_t Dr = eq_mask16(vr, 0x2e2e2e2eu) & V, Dc = eq_mask16(vc, 0x2e2e2e2eu) & V, Du = eq_mask16(vu, 0x2e2e2e2eu) & V;
        const uint32_t Erc = eq_pair16(vr, vc), Eru = eq_pair16(vr, vu), Euc = eq_pair16(vu, vc);
        acc[13] += __popc(gc_mask16(vr) & V); acc[14] += __popc(gc_mask16(vc) & V);
        acc[15] += __popc(Dr); acc[16] += __popc(Dc); acc[17] += __popc(Du);
        uint32_t OK = V & range_mask16(cm.lmask - i0, cm.rmask - i0);
        for (int k = 0; k < cm.nkeys; ++k) {
          const uint32_t in = V & range_mask16(cm.key_a[k] - i0, cm.key_b[k] - i0);
          OK &= ~in;
          acc[18] += __popc(Dr & in);                       // dots of the reference row inside stretches
        }
        const uint32_t Nrc = OK & ~Erc, Nru = OK & ~Eru;
        acc[7] += __popc(Nrc & Dr); acc[9] += __popc(Nrc & ~Dr & ~Dc); acc[8] += __popc(Nrc & ~Dr & Dc);
        acc[10] += __popc(Nru & Dr); acc[12] += __popc(Nru & ~Dr & ~Du); acc[11] += __popc(Nru & ~Dr & Du);
        const uint32_t A = OK & Eru, X = A & ~Euc, Y = A & Euc;
        const uint32_t B1 = Nru & Erc, B2 = Nru & ~Erc, B3 = B2 & Euc;
        acc[5] += __popc(A); acc[6] += __popc(Nru);
        acc[0] += __popc(Y) + __popc(B1); acc[3] += __popc(Y) + __popc(B1);
        acc[1] += __popc(X) + __popc(B3); acc[4] += __popc(X) + __popc(B2);
        acc[2] += __popc(B3);
      }
    } else {
      for (int i = threadIdx.x; i < L; i += blockDim.x) classify_column(i, rr[i], cc[i], uu[i], cm, acc);
    }
  }
  TT(2)
#pragma unroll
  for (int k = 0; k < kNAcc; ++k) {
    int v = acc[k];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sm[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t[kNAcc];
    for (int k = 0; k < kNAcc; ++k) t[k] = sm[0][k] + sm[1][k] + sm[2][k] + sm[3][k];
    int64_t *o = counters + r * ELECTOR_TALLY_K;
    for (int k = 0; k < ELECTOR_TALLY_K; ++k) o[k] = 0;
    o[ELECTOR_T_NCOLS] = L;
    o[ELECTOR_T_EXTENDED] = -1;
    if (assessed) {
      for (int k = 0; k < 15; ++k) o[k] = t[k];
      o[ELECTOR_T_LENREF] = L - t[15]; o[ELECTOR_T_LENCOR] = L - t[16]; o[ELECTOR_T_LENUNC] = L - t[17];
      o[ELECTOR_T_GAPSLEFT] = sc.gl; o[ELECTOR_T_GAPSRIGHT] = sc.gr;
      int64_t missing = 0;
      for (int k = 0; k < sc.nkeys; ++k) missing += sc.key_b[k] - sc.key_a[k];
      missing -= t[18];
      missing -= sc.gl + sc.gr;
      o[ELECTOR_T_MISSING] = missing < 0 ? 0 : missing;
      o[ELECTOR_T_EXTENDED] = sc.ext;
      o[ELECTOR_T_ASSESSED] = 1;
    }
    if (stretches) {   // the border gap stretches (findGapStretches' dict): what the report needs to rebuild the column mask
      int32_t *q = stretches + r * ELECTOR_STRETCH_K;
      q[0] = assessed ? sc.nkeys : 0;
      for (int k = 0; k < kMaxStretchKeys; ++k) { q[1 + 2 * k] = sc.key_a[k]; q[2 + 2 * k] = sc.key_b[k]; }
    }
  }
  TT(3)
}

// ---- global counters: sums[k] += sum over reads of counters[r][k] (the "final counter reduction") ----
__global__ void __launch_bounds__(256) tally_sum_kernel(int64_t n_reads, const int64_t *counters, unsigned long long *sums) {
  __shared__ unsigned long long part[ELECTOR_TALLY_K];
  if (threadIdx.x < ELECTOR_TALLY_K) part[threadIdx.x] = 0;
  __syncthreads();
  const int64_t total = n_reads * ELECTOR_TALLY_K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % ELECTOR_TALLY_K);
    const int64_t v = counters[i];
    // "extended" is -1 for reads that were not extended: only extended reads count
    if (k != ELECTOR_T_EXTENDED || v >= 0) atomicAdd(&part[k], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < ELECTOR_TALLY_K && part[threadIdx.x]) atomicAdd(&sums[threadIdx.x], part[threadIdx.x]);
}

}  // namespace elector

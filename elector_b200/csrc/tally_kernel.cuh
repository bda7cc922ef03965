// tally_kernel.cuh -- device merge (Donatello semantics) + per-read tally (computeStats.py).
//
//   read_totals_kernel / scan_offsets_kernel  : where each read's merged rows start
//   merge_rows_kernel  (Donatello.cpp:13-31,50-84): concatenate a read's window MSAs, dropping
//                       every column whose corrected row is 'n'; one warp per read
//   tally_scan_kernel  (computeStats.py:61-98,104-189,472-498): the sequential scanners --
//                       left/right gaps, extension, gap stretches; one thread per read
//   tally_count_kernel (computeStats.py:291-328,371-440,712-752): per-column classification
//                       under the "existing corrected positions" mask; one CTA per read,
//                       warp-shuffle reduction of the counters
// All three are byte streaming kernels (HBM-bound, 3 bytes per MSA column).
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
__global__ void read_totals_kernel(int64_t n_reads, const int64_t *read_first, const int32_t *nring, int64_t *tot) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
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
__global__ void __launch_bounds__(128) merge_rows_kernel(int64_t n_reads, const int64_t *read_first, const uint8_t *rows,
                                                          const int64_t *row_off, const int32_t *row_stride, const int32_t *nring,
                                                          const int64_t *m_off, uint8_t *m_ref, uint8_t *m_cor, uint8_t *m_unc,
                                                          int32_t *m_len) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_reads) return;
  const int64_t w0 = read_first[r], w1 = read_first[r + 1];
  const int64_t base = m_off[r];
  int64_t done = 0;
  for (int64_t wb = w0; wb < w1; wb += 32) {
    const int64_t w = wb + lane;
    int kept = 0, k = 0, st = 0;
    const uint8_t *src = nullptr;
    if (w < w1) {
      k = nring[w]; st = row_stride[w]; src = rows + row_off[w];
      for (int i = 0; i < k; ++i) kept += (src[st + i] != 'n');
    }
    int incl = kept;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (w < w1) {
      int64_t o = base + done + incl - kept;
      for (int i = 0; i < k; ++i) {
        const uint8_t c = src[st + i];
        if (c != 'n') { m_ref[o] = src[i]; m_cor[o] = c; m_unc[o] = src[2 * st + i]; ++o; }
      }
    }
    done += total;
  }
  if (lane == 0) m_len[r] = (int32_t)done;
}

// ---- sequential scanners ---------------------------------------------------------------
__device__ __forceinline__ int nb_left_gaps(const uint8_t *s, int L) {
  int gaps = 0, nt = 0, total = 0;
  for (int i = 0; i < L && nt <= T_THRESH; ++i) {
    if (s[i] == '.') { ++gaps; nt = 0; }
    else { if (gaps >= T_THRESH) total = i; gaps = 0; ++nt; }
  }
  return total;
}
__device__ __forceinline__ int nb_right_gaps(const uint8_t *s, int L) {
  int gaps = 0, nt = 0, total = 0;
  for (int i = L - 1; i >= 0 && nt <= T_THRESH; --i) {
    if (s[i] == '.') { ++gaps; nt = 0; }
    else { if (gaps >= T_THRESH) total = L - i; gaps = 0; ++nt; }
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

__global__ void __launch_bounds__(128) tally_scan_kernel(int64_t n_reads, const uint8_t *R, const uint8_t *C, const uint8_t *U,
                                                          const int64_t *off, const int32_t *len, ReadScan *out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int L = len ? len[r] : (int)(off[r + 1] - off[r]);
  const uint8_t *rr = R + off[r], *cc = C + off[r], *uu = U + off[r];
  ReadScan o;
  o.gl = o.gr = 0; o.ext = -1; o.nkeys = 0; o.overflow = 0;
  for (int k = 0; k < kMaxStretchKeys; ++k) o.key_a[k] = o.key_b[k] = 0;
  if (L > 10) {
    o.gl = min(nb_left_gaps(rr, L), nb_left_gaps(uu, L));
    o.gr = min(nb_right_gaps(rr, L), nb_right_gaps(uu, L));
    int ext = -1;
    if (o.gl >= T_THRESH2) { int dots = 0; for (int i = 0; i < o.gl; ++i) dots += cc[i] == '.'; ext = (ext < 0 ? 0 : ext) + o.gl - dots; }
    if (o.gr >= T_THRESH2) { int dots = 0; for (int i = L - o.gr + 1; i < L; ++i) dots += cc[i] == '.'; ext = (ext < 0 ? 0 : ext) + o.gr - dots; }
    o.ext = ext;
    // findGapStretches scan (:111-142), slots consumed as soon as they are final
    StretchState s;
    s.L = L; s.has0 = s.end0 = s.nkeys = s.overflow = s.pend = s.pa = s.pb = s.merge = 0;
    int nslots = 0, cg = 0, cr = 0, ca = 0, cb = 0;
    bool have_cur = false, cur_set = false, prev_dot = false;
    for (int pos = 0; pos < L; ++pos) {
      const bool cd = cc[pos] == '.', rd = rr[pos] == '.';
      if (pos == 0) { cg += cd; cr += rd; }
      else if (prev_dot) {
        if (cd) cg = cg > 0 ? cg + 1 : 2;
        if (rd) cr = cr > 0 ? cr + 1 : 2;
      }
      if (!cd) {
        if (cg > 0) {
          if (have_cur && cur_set) s.finalize(ca, cb, true);
          have_cur = true; cur_set = false; ++nslots;
        }
        cg = 0;
      }
      if (!rd) cr = 0;
      if (cg >= T_THRESH && cr < T_THRESH2) {
        if (nslots == 0) { have_cur = true; cur_set = true; nslots = 1; ca = pos - T_THRESH + 1; cb = pos; }
        else { if (!cur_set) { cur_set = true; ca = pos - T_THRESH + 1; } cb = pos; }
      }
      prev_dot = cd;
    }
    if (have_cur && cur_set) s.finalize(ca, cb, nslots > 1);
    if (s.pend && !s.merge) s.emit2(s.pa, s.pb);
    if (s.has0) { o.key_a[o.nkeys] = 0; o.key_b[o.nkeys] = s.end0; ++o.nkeys; }
    for (int k = 0; k < s.nkeys && o.nkeys < kMaxStretchKeys; ++k) { o.key_a[o.nkeys] = s.key[k]; o.key_b[o.nkeys] = L - 1; ++o.nkeys; }
    o.overflow = s.overflow || (s.has0 + s.nkeys > kMaxStretchKeys);
  }
  out[r] = o;
}

// ---- per-column counters ---------------------------------------------------------------
constexpr int kNAcc = 19;  // accumulators reduced per read (see below)

__global__ void __launch_bounds__(128) tally_count_kernel(int64_t n_reads, const uint8_t *R, const uint8_t *C, const uint8_t *U,
                                                           const int64_t *off, const int32_t *len, const ReadScan *scan,
                                                           int64_t *counters) {
  const int64_t r = blockIdx.x;
  if (r >= n_reads) return;
  const int L = len ? len[r] : (int)(off[r + 1] - off[r]);
  const uint8_t *rr = R + off[r], *cc = C + off[r], *uu = U + off[r];
  const ReadScan sc = scan[r];
  int acc[kNAcc];
#pragma unroll
  for (int k = 0; k < kNAcc; ++k) acc[k] = 0;
  const bool assessed = L > 10;
  const int lmask = sc.gl >= T_THRESH ? sc.gl : 0;                 // columns [0, gl) masked
  const int rmask = sc.gr >= T_THRESH ? L - sc.gr : L - 1;         // columns (L-gr, L-1] masked
  if (assessed) {
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const uint8_t r_ = rr[i], c_ = cc[i], u_ = uu[i];
      acc[13] += (r_ == 'g' || r_ == 'c' || r_ == 'G' || r_ == 'C');
      acc[14] += (c_ == 'g' || c_ == 'c' || c_ == 'G' || c_ == 'C');
      acc[15] += r_ == '.'; acc[16] += c_ == '.'; acc[17] += u_ == '.';
      bool ok = i >= lmask && i <= rmask;
      for (int k = 0; k < sc.nkeys; ++k) {
        const bool in = i >= sc.key_a[k] && i <= sc.key_b[k];
        if (in) { ok = false; acc[18] += (r_ == '.'); }           // dots of the reference row inside stretches
      }
      if (!ok) continue;
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
  }
  __shared__ int sm[4][kNAcc];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
  }
}

}  // namespace elector

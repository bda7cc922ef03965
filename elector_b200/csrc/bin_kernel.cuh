// bin_kernel.cuh -- device-side scheduling of the windows of one call.
//
// The POA kernel runs one window per thread in lock step, so a warp is only efficient when
// its 32 windows have the same loop shape: the same number of 8-row bands in both DPs and a
// similar number of columns.  These three small kernels sort the window ids by
// (bands of DP2, bands of DP1, reference length / 4), largest first, with a counting sort
// (histogram -> single-CTA scan -> scatter), and cut the sorted list into a few SEGMENTS
// (one launch each) whose per-warp scratch is sized by the segment's own maxima.
// Everything stays on the device; the host reads back one 1 KB table.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elector {

constexpr int kBigTiers = 8;          // windows with a sequence longer than 256: one bin per power of two
constexpr int kSmallMax = 256;        // longest sequence of a "small" window
constexpr int kNbMax = kSmallMax / 8; // bands of a small window: 1..32
constexpr int kXq = 65;               // reference-length quanta (lr / 4: 0..64)
constexpr int kSmallBins = kNbMax * kNbMax * kXq;
constexpr int kNumBins = kBigTiers + kSmallBins;
constexpr int kNumSegs = kBigTiers + 4;
constexpr int kMaxWindowLen = 32000;  // node records index their ordinal slot with 15 bits

struct SegInfo {      // one launch of the POA kernel
  int32_t count;      // windows
  int32_t start;      // first position in the sorted item list
  int32_t max_lr, max_lc, max_lu;
  int32_t pad[3];
};

struct BinTable {
  SegInfo seg[kNumSegs];
  int32_t err_code;    // 0 ok, 1 empty sequence, 2 sequence longer than kMaxWindowLen
  int32_t err_window;  // smallest offending window id
};

__device__ __forceinline__ int seg_of_small(int nb2) { return kBigTiers + (nb2 > 16 ? 0 : nb2 > 8 ? 1 : nb2 > 4 ? 2 : 3); }

__device__ __forceinline__ void bin_of(int lr, int lc, int lu, int &bin, int &seg) {
  const int mx = max(lr, max(lc, lu));
  if (mx > kSmallMax) {
    int t = 1;
    while ((kSmallMax << t) < mx) ++t;  // t = 1: <= 512, 2: <= 1024, ...
    t = min(t, kBigTiers);
    bin = seg = kBigTiers - t;          // the largest tier comes first
  } else {
    const int nb2 = (lu + 7) >> 3, nb1 = (lc + 7) >> 3, xq = lr >> 2;
    const int small = ((nb2 - 1) * kNbMax + (nb1 - 1)) * kXq + xq;
    bin = kBigTiers + (kSmallBins - 1 - small);
    seg = seg_of_small(nb2);
  }
}

// hist[bin] += 1; per-segment maxima; validation
__global__ void __launch_bounds__(256) bin_count_kernel(int32_t n, const int64_t *ro, const int64_t *co, const int64_t *uo,
                                                         int32_t *hist, BinTable *tab) {
  for (int32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const int64_t lr64 = ro[w + 1] - ro[w], lc64 = co[w + 1] - co[w], lu64 = uo[w + 1] - uo[w];
    if (lr64 <= 0 || lc64 <= 0 || lu64 <= 0) { atomicMax(&tab->err_code, 1); atomicMin(&tab->err_window, w); continue; }
    if (lr64 > kMaxWindowLen || lc64 > kMaxWindowLen || lu64 > kMaxWindowLen) { atomicMax(&tab->err_code, 2); atomicMin(&tab->err_window, w); continue; }
    const int lr = (int)lr64, lc = (int)lc64, lu = (int)lu64;
    int bin, seg;
    bin_of(lr, lc, lu, bin, seg);
    atomicAdd(&hist[bin], 1);
    SegInfo *s = &tab->seg[seg];
    if (lr > s->max_lr) atomicMax(&s->max_lr, lr);
    if (lc > s->max_lc) atomicMax(&s->max_lc, lc);
    if (lu > s->max_lu) atomicMax(&s->max_lu, lu);
  }
}

// exclusive scan of the histogram in place (hist[bin] becomes the bin's first position) and
// the start / count of every segment; one CTA (67 608 bins: a few microseconds)
__global__ void __launch_bounds__(1024) bin_scan_kernel(int32_t *hist, BinTable *tab) {
  __shared__ int32_t part[1024];
  constexpr int per = (kNumBins + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, kNumBins);
  int32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += hist[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {  // Hillis-Steele inclusive scan of the partial sums
    const int32_t t = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  int32_t run = part[threadIdx.x] - sum;
  for (int i = lo; i < hi; ++i) { const int32_t c = hist[i]; hist[i] = run; run += c; }
  __syncthreads();
  if (threadIdx.x < kNumSegs) {
    // first bin of every segment (bins are ordered largest first; small bins descend in nb2)
    auto first_bin = [](int seg) {
      if (seg < kBigTiers) return seg;
      if (seg >= kNumSegs) return kNumBins;
      const int nb2_hi = seg == kBigTiers ? 32 : seg == kBigTiers + 1 ? 16 : seg == kBigTiers + 2 ? 8 : 4;
      const int small_hi = ((nb2_hi - 1) * kNbMax + (kNbMax - 1)) * kXq + (kXq - 1);
      return kBigTiers + (kSmallBins - 1 - small_hi);
    };
    const int s = threadIdx.x;
    const int b0 = first_bin(s), b1 = first_bin(s + 1);
    const int32_t total = part[1023];
    const int32_t p0 = b0 < kNumBins ? hist[b0] : total, p1 = b1 < kNumBins ? hist[b1] : total;
    tab->seg[s].start = p0;
    tab->seg[s].count = p1 - p0;
  }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(int32_t n, const int64_t *ro, const int64_t *co, const int64_t *uo,
                                                           int32_t *cursor, int32_t *items) {
  for (int32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const int64_t lr = ro[w + 1] - ro[w], lc = co[w + 1] - co[w], lu = uo[w + 1] - uo[w];
    if (lr <= 0 || lc <= 0 || lu <= 0 || lr > kMaxWindowLen || lc > kMaxWindowLen || lu > kMaxWindowLen) continue;  // reported by bin_count_kernel
    int bin, seg;
    bin_of((int)lr, (int)lc, (int)lu, bin, seg);
    items[atomicAdd(&cursor[bin], 1)] = w;
  }
}

}  // namespace elector

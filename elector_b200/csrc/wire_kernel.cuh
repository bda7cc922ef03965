// wire_kernel.cuh -- the compact wire format of elector_pipeline_run2 (include/elector_poa.h): 2-bit letters and 32-bit
// offsets in, 4-bit merged columns out.  The alignment kernels keep reading one byte per letter and 64-bit offsets: these
// kernels expand / pack on the device, where a pass over the letters of a call costs ~0.1 ms of HBM time -- the host link
// carries a quarter of the bytes.  All are streaming kernels (HBM-bound).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elector {

// bits: the packed letters from byte (first >> 2) of the call's stream on; out[i] = letter first + i, i < n
__global__ void __launch_bounds__(256) unpack2_kernel(const uint8_t *bits, int shift, int64_t n, uint8_t *out) {
  const uint32_t lut = 0x54474341u;   // 'A' 'C' 'G' 'T'
  const int64_t nw = (n + 3) >> 2;    // output words
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nw; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = 4 * t + shift;  // first letter of the word, counted from bits[0]
    const uint32_t two = (uint32_t)bits[g >> 2] | ((uint32_t)bits[(g >> 2) + 1] << 8);   // (the caller pads the copy by one byte)
    const uint32_t v = two >> (2 * (g & 3));
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= ((lut >> (8 * ((v >> (2 * k)) & 3u))) & 0xffu) << (8 * k);
    reinterpret_cast<uint32_t *>(out)[t] = w;   // out is 4-byte aligned and padded to a multiple of 4
  }
}

// the letters that are not A, C, G, T: pos = ascending positions in the call's stream, first = position of out[0]
__global__ void patch_letters_kernel(int64_t n_exc, const int64_t *pos, const uint8_t *byte, int64_t first, uint8_t *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_exc) out[pos[i] - first] = byte[i];
}

// 32-bit offsets relative to the chunk's first letter -> the 64-bit offsets of the whole call the kernels index with
__global__ void widen_offsets_kernel(int64_t n, const int32_t *rel, int64_t base, int64_t *off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] = base + rel[i];
}

// 16-bit window lengths -> the 32-bit ones the scan below works on
__global__ void widen_len16_kernel(int64_t n, const uint16_t *len16, int32_t *len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = len16[i];
}

// window lengths -> 64-bit offsets: per 1024-window block an exclusive scan (in place: len becomes the offset inside the
// block) and the block total; then one CTA scans the block totals (at most 1024 blocks x 1024 windows per chunk); then
// off[i] = base + block_base[i / 1024] + len[i], off[n] = base + total
__global__ void __launch_bounds__(1024) len_scan_blocks_kernel(int64_t n, int32_t *len, long long *block_total) {
  __shared__ int32_t warp_sum[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t v = i < n ? len[i] : 0;
  int32_t incl = v;
  for (int d = 1; d < 32; d <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int32_t s = warp_sum[lane];
    for (int d = 1; d < 32; d <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += t; }
    warp_sum[lane] = s;
  }
  __syncthreads();
  const int32_t before = (wid ? warp_sum[wid - 1] : 0) + incl - v;
  if (i < n) len[i] = before;
  if (threadIdx.x == 1023) block_total[blockIdx.x] = before + v;
}
__global__ void __launch_bounds__(1024) len_scan_totals_kernel(int nblocks, long long *block_total) {
  __shared__ long long part[1024];
  const long long v = (int)threadIdx.x < nblocks ? block_total[threadIdx.x] : 0;
  part[threadIdx.x] = v;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const long long t = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  if ((int)threadIdx.x < nblocks) block_total[threadIdx.x] = part[threadIdx.x] - v;
  if (threadIdx.x == 1023) block_total[nblocks] = part[1023];
}
__global__ void len_to_offsets_kernel(int64_t n, const int32_t *scanned, const long long *block_base, int nblocks, int64_t base, int64_t *off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] = base + block_base[i >> 10] + scanned[i];
  else if (i == n) off[n] = base + block_base[nblocks];
}

// merged rows of the reads -> two columns per byte (ELECTOR_NIBBLE_CHARS; code 15 + an entry in the escape list for any
// other character).  One warp per (read, row); col_base = column of the chunk's first column in the caller's buffers.
__global__ void __launch_bounds__(128) nibble_pack_kernel(int64_t n_reads, const uint8_t *m0, const uint8_t *m1, const uint8_t *m2,
                                                           const int64_t *m_off, const int32_t *m_len, uint8_t *n0, uint8_t *n1, uint8_t *n2,
                                                           int64_t col_base, unsigned long long *esc_count, int64_t *esc_pos, uint8_t *esc_byte,
                                                           int64_t esc_cap, const int32_t *abort) {
  if (*abort) return;
  const int lane = threadIdx.x & 31;
  const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= 3 * n_reads) return;
  const int64_t r = job / 3;
  const int row = (int)(job - 3 * r);
  const uint8_t *src = (row == 0 ? m0 : row == 1 ? m1 : m2) + m_off[r];
  uint8_t *dst = (row == 0 ? n0 : row == 1 ? n1 : n2) + (m_off[r] >> 1);
  const int L = m_len[r];
  auto code = [&](int i) -> uint32_t {
    if (i >= L) return 0;
    const uint8_t c = src[i];
    const uint32_t k = c == '.' ? 0u : c == 'a' ? 1u : c == 'c' ? 2u : c == 'g' ? 3u : c == 't' ? 4u : c == 'n' ? 5u : c == 'A' ? 6u : 15u;
    if (k == 15u) {
      const unsigned long long e = atomicAdd(esc_count, 1ull);
      if ((int64_t)e < esc_cap) { esc_pos[e] = 3 * (col_base + m_off[r] + i) + row; esc_byte[e] = c; }
    }
    return k;
  };
  for (int i = 2 * lane; i < L; i += 64) dst[i >> 1] = (uint8_t)(code(i) | (code(i + 1) << 4));
}

// merged rows of the reads -> ONE byte per column for the three rows together (m_nibbles = 2): ref + 6 * cor + 36 * unc with
// each row's character coded over ELECTOR_COLUMN_CHARS; a column with any other character gets 255 and its three
// characters go to the escape list.  A thread packs four columns (read regions start at multiples of 16 columns);
// one warp per 128 columns of a read.
__device__ __forceinline__ uint32_t column_code6(uint32_t c) {
  return c == '.' ? 0u : c == 'a' ? 1u : c == 'c' ? 2u : c == 'g' ? 3u : c == 't' ? 4u : c == 'n' ? 5u : 15u;
}
__global__ void __launch_bounds__(128) column_pack_kernel(int64_t n_reads, const uint8_t *m0, const uint8_t *m1, const uint8_t *m2,
                                                           const int64_t *m_off, const int32_t *m_len, uint8_t *out, int64_t col_base,
                                                           unsigned long long *esc_count, int64_t *esc_pos, uint8_t *esc_byte, int64_t esc_cap,
                                                           const int32_t *abort) {
  if (*abort) return;
  const int64_t r = blockIdx.x;
  if (r >= n_reads) return;
  const int64_t off = m_off[r];
  const int L = m_len[r];
  const uint32_t *a4 = reinterpret_cast<const uint32_t *>(m0 + off), *c4 = reinterpret_cast<const uint32_t *>(m1 + off),
                 *u4 = reinterpret_cast<const uint32_t *>(m2 + off);
  uint32_t *o4 = reinterpret_cast<uint32_t *>(out + off);
  for (int q = threadIdx.x; 4 * q < L; q += blockDim.x) {
    const uint32_t a = a4[q], c = c4[q], u = u4[q];
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * q + k;
      if (i < L) {
        const uint32_t ca = (a >> (8 * k)) & 0xffu, cc = (c >> (8 * k)) & 0xffu, cu = (u >> (8 * k)) & 0xffu;
        const uint32_t ka = column_code6(ca), kc = column_code6(cc), ku = column_code6(cu);
        uint32_t code = ka + 6u * kc + 36u * ku;
        if ((ka | kc | ku) == 15u) {
          code = 255u;
          const unsigned long long e = atomicAdd(esc_count, 3ull);
          if ((int64_t)e + 3 <= esc_cap) {
            const int64_t col = col_base + off + i;
            esc_pos[e] = 3 * col; esc_byte[e] = (uint8_t)ca;
            esc_pos[e + 1] = 3 * col + 1; esc_byte[e + 1] = (uint8_t)cc;
            esc_pos[e + 2] = 3 * col + 2; esc_byte[e + 2] = (uint8_t)cu;
          }
        }
        w |= code << (8 * k);
      }
    }
    o4[q] = w;
  }
}

}  // namespace elector

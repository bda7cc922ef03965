// split_kernel.cuh -- ELECTOR's window cutting (src/split/Master_Splitter.cpp) on the device: SURVEY.md 8(f)-1.
//
// The reference cuts every (reference, uncorrected, corrected) read triplet into ~50-letter windows at k-mers that occur
// exactly once in each of the three reads (split() :175-308), keeps the cutting with the smallest largest window over
// k = 15, 13, 11, 9 (best_split() :310-332), and hands the windows to `poa`.  It is a serial 10.7 ms per triplet; with the
// alignment at 0.7 us per triplet it is 99.9 % of the stage.
//
// Mapping: ONE CTA PER (triplet, k) JOB -- the four k of a triplet run side by side, the choice between them is made
// afterwards.  Per job:
//   1. the k-mers of the reference go into an open-addressing hash table in global memory (L2-resident: 20 bytes per slot,
//      2 slots per k-mer); occurrence flags "seen once" / "seen again" per read are set with atomicOr, so "exactly once in
//      each read" is a flag pattern (:176-230 keep three std::unordered_maps with -1 for repeats);
//   2. the k-mers of the other two reads look their slot up and set their flags and positions;
//   3. every reference position asks the table whether its k-mer is such an anchor -> a bitmap in reference order;
//   4. one thread thins the bitmap like the reference's left-to-right scan (:242-251, an anchor at most every minSize + 1
//      letters), the CTA fetches the anchors' positions in the other two reads;
//   5. the longest chain of anchors that increase by less than 1000 letters in all three reads (:79-126, a memoised
//      recursion) is a backward dynamic programme over the anchor list: a warp looks at the <= 48 successors of an anchor
//      at once; ties go to the first, like the reference's strict comparisons;
//   6. one thread walks the chain and cuts (:262-306), including the two special cases of a corrected read that starts late
//      or ends early (the reference then splits reference and uncorrected read alone and pads the corrected side with `N`
//      records, :268-277 and :295-301): the CTA runs steps 1-5 again on the sub-strings.
// The same code compiles for the host (tests/emul/split_emul.cu runs it serially against the compiled reference).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define SP_HD __host__ __device__
#else
#define SP_HD
#endif

namespace elector {

constexpr uint32_t kSplitEmpty = 0xffffffffu;
enum : uint32_t { SF_R1 = 1, SF_R2 = 2, SF_A1 = 4, SF_A2 = 8, SF_B1 = 16, SF_B2 = 32 };   // seen once / again in ref, S1, S2
constexpr int kChainReach = 1000;   // Master_Splitter.cpp:86-88

struct SplitSeq { const uint8_t *s; int n; };

// one output record of a job: three (start, length) pairs into the job's three reads; b0 = -1: the corrected side is the
// placeholder "N" (generate_dumb_str, :139-154)
struct SplitWin { int32_t r0, rn, a0, an, b0, bn; };

// scratch of one job (global memory; sized for the longest read of the call)
struct SplitScratch {
  uint32_t *key, *flag, *posr, *posa, *posb;   // hash table, `slots` entries each
  uint32_t slots;                               // power of two
  uint32_t *cand;                               // bitmap over reference positions
  int32_t *ar, *aa, *ab, *chain, *nxt;          // anchors (positions in the three reads), chain length from here, successor
  int32_t max_anchors;
  int32_t *bl;                                  // the chain: indices into the anchor list
};

// letter codes of the reference's two encoders: str2num (:26-38) for the first k letters of a read, nuc2int (:41-49) for the rest
SP_HD inline uint32_t split_code(const SplitSeq &q, int t, int k) {
  const uint8_t c = q.s[t];
  if (t < k) return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;
  return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}
// the k-mer that the reference has in `seq` when it is at position p of the read (positions 0 .. kmers(q) - 1)
SP_HD inline uint32_t split_kmer(const SplitSeq &q, int p, int k) {
  uint32_t v = 0;
  const int e = p + k < q.n ? p + k : q.n;     // a read shorter than k has one, shorter, k-mer (substr(0, k), :177)
  for (int t = p; t < e; ++t) v = (v << 2) | split_code(q, t, k);
  return v & ((1u << (2 * k)) - 1u);
}
SP_HD inline int split_kmers(const SplitSeq &q, int k) { return q.n > k ? q.n - k + 1 : 1; }
SP_HD inline uint32_t split_hash(uint32_t kmer) { kmer *= 0x9E3779B1u; return kmer ^ (kmer >> 15); }

#ifdef __CUDA_ARCH__
#define SP_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define SP_SYNC() __syncthreads()
#define SP_SERIAL if (threadIdx.x == 0)
SP_HD inline uint32_t sp_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
SP_HD inline uint32_t sp_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
#else
#define SP_FOR(i, n) for (int i = 0; i < (n); ++i)
#define SP_SYNC()
#define SP_SERIAL
SP_HD inline uint32_t sp_cas(uint32_t *p, uint32_t cmp, uint32_t v) { const uint32_t o = *p; if (o == cmp) *p = v; return o; }
SP_HD inline uint32_t sp_or(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o | v; return o; }
#endif

// slot of `kmer`, inserting it when `insert`; kSplitEmpty when absent
SP_HD inline uint32_t split_slot(const SplitScratch &sc, uint32_t kmer, bool insert) {
  const uint32_t mask = sc.slots - 1;
  uint32_t s = split_hash(kmer) & mask;
  for (;;) {
    uint32_t cur = sc.key[s];
    if (cur == kSplitEmpty) {
      if (!insert) return kSplitEmpty;
      cur = sp_cas(&sc.key[s], kSplitEmpty, kmer);
      if (cur == kSplitEmpty) return s;
    }
    if (cur == kmer) return s;
    s = (s + 1) & mask;
  }
}

// Steps 1-5 for (ref, S1, S2): the anchor list and its best chain.  Returns the chain length (0: no anchor at all) in every
// thread; the chain is sc.bl[0 .. len-1].  s_int: four ints of shared scratch (device) / any four ints (host).
SP_HD inline int split_chain(const SplitScratch &sc, const SplitSeq &ref, const SplitSeq &S1, const SplitSeq &S2, int k, uint32_t min_size, int *s_int) {
  // 1. table of the reference k-mers
  SP_FOR(i, (int)sc.slots) { sc.key[i] = kSplitEmpty; sc.flag[i] = 0; }
  SP_FOR(i, (ref.n + 31) / 32 + 1) sc.cand[i] = 0;
  SP_SYNC();
  const int nr = split_kmers(ref, k), na = split_kmers(S1, k), nb = split_kmers(S2, k);
  SP_FOR(p, nr) {
    const uint32_t s = split_slot(sc, split_kmer(ref, p, k), true);
    if (sp_or(&sc.flag[s], SF_R1) & SF_R1) sp_or(&sc.flag[s], SF_R2);
    sc.posr[s] = (uint32_t)p;   // used only when the k-mer occurs once
  }
  SP_SYNC();
  // 2. the other two reads
  SP_FOR(p, na) {
    const uint32_t s = split_slot(sc, split_kmer(S1, p, k), false);
    if (s != kSplitEmpty) { if (sp_or(&sc.flag[s], SF_A1) & SF_A1) sp_or(&sc.flag[s], SF_A2); sc.posa[s] = (uint32_t)p; }
  }
  SP_FOR(p, nb) {
    const uint32_t s = split_slot(sc, split_kmer(S2, p, k), false);
    if (s != kSplitEmpty) { if (sp_or(&sc.flag[s], SF_B1) & SF_B1) sp_or(&sc.flag[s], SF_B2); sc.posb[s] = (uint32_t)p; }
  }
  SP_SYNC();
  // 3. anchors in reference order: once in each read
  SP_FOR(p, nr) {
    const uint32_t s = split_slot(sc, split_kmer(ref, p, k), false);
    if (sc.flag[s] == (SF_R1 | SF_A1 | SF_B1)) sp_or(&sc.cand[p >> 5], 1u << (p & 31));
  }
  SP_SYNC();
  // 4. left-to-right thinning (:242-251): position 0 is taken as it is; position j + 1 when j - last > minSize, last = j
  SP_SERIAL {
    int n = 0;
    uint32_t last = 0;
    const int words = (nr + 31) / 32;
    for (int w = 0; w < words; ++w) {
      uint32_t bits = sc.cand[w];
      while (bits) {
        int b = 0;
        while (!((bits >> b) & 1u)) ++b;
        bits &= bits - 1;
        const int p = w * 32 + b;
        if (p == 0) { if (n < sc.max_anchors) sc.ar[n++] = 0; continue; }
        const uint32_t j = (uint32_t)(p - 1);
        if (j - last > min_size) { if (n < sc.max_anchors) sc.ar[n++] = p; last = j; }
      }
    }
    s_int[0] = n;
  }
  SP_SYNC();
  const int n = s_int[0];
  SP_FOR(i, n) {
    const uint32_t s = split_slot(sc, split_kmer(ref, sc.ar[i], k), false);
    sc.aa[i] = (int32_t)sc.posa[s];
    sc.ab[i] = (int32_t)sc.posb[s];
  }
  SP_SYNC();
  // 5. longest chain, backwards: chain[i] = 1 + max over the successors within reach (first maximum), 0 when there is none
#ifdef __CUDA_ARCH__
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int i = n - 1; i >= 0; --i) {
      const int r = sc.ar[i], a = sc.aa[i], b = sc.ab[i];
      int best = -1, arg = -1;
      for (int base = i + 1; base < n; base += 32) {          // successors in list order until the reference distance reaches 1000 (:86,:99)
        const int c = base + lane;
        bool in = c < n && sc.ar[c] - r < kChainReach;
        if (in) {
          const int da = sc.aa[c] - a, db = sc.ab[c] - b;
          if (da > 0 && da < kChainReach && db > 0 && db < kChainReach) { const int v = sc.chain[c]; if (v > best) { best = v; arg = c; } }
        }
        const bool stop = __any_sync(0xffffffffu, c < n && !in) || base + 32 >= n;
        if (stop) break;
      }
      for (int d = 16; d > 0; d >>= 1) {                       // maximum, ties to the smaller index
        const int ob = __shfl_xor_sync(0xffffffffu, best, d), oa = __shfl_xor_sync(0xffffffffu, arg, d);
        if (ob > best || (ob == best && oa >= 0 && (arg < 0 || oa < arg))) { best = ob; arg = oa; }
      }
      if (lane == 0) { sc.chain[i] = 1 + best; sc.nxt[i] = arg; }
      __syncwarp();
    }
  }
#else
  for (int i = n - 1; i >= 0; --i) {
    int best = -1, arg = -1;
    for (int c = i + 1; c < n; ++c) {
      if (!(sc.ar[c] - sc.ar[i] < kChainReach)) break;
      const int da = sc.aa[c] - sc.aa[i], db = sc.ab[c] - sc.ab[i];
      if (da > 0 && da < kChainReach && db > 0 && db < kChainReach && sc.chain[c] > best) { best = sc.chain[c]; arg = c; }
    }
    sc.chain[i] = 1 + best; sc.nxt[i] = arg;
  }
#endif
  SP_SYNC();
  SP_SERIAL {   // the first anchor with the longest chain starts it (:113-119), then the successors
    int best = -1, at = -1;
    for (int i = 0; i < n; ++i) if (sc.chain[i] > best) { best = sc.chain[i]; at = i; }
    int len = 0;
    while (at != -1) { sc.bl[len++] = at; at = sc.nxt[at]; }
    s_int[1] = len;
  }
  SP_SYNC();
  return s_int[1];
}

// The cutting of one job (split(), :175-308) into out[0 ..); returns the number of records, or -1 when out_cap is too small.
// Positions in the records are relative to the three reads given here.  first_call as in the reference: the late-start /
// early-end special cases only at the outer level.
SP_HD inline int split_job(const SplitScratch &sc, const SplitScratch &sc2, SplitSeq ref, SplitSeq S1, SplitSeq S2, int k, SplitWin *out, int out_cap, int *s_int) {
  const uint32_t min_size = 20;
  const int blen = split_chain(sc, ref, S1, S2, k, min_size, s_int);
  if (blen < 1) {   // no anchor: the three reads as one record (:256-261)
    SP_SERIAL { if (out_cap >= 1) out[0] = SplitWin{0, ref.n, 0, S1.n, 0, S2.n}; s_int[2] = out_cap >= 1 ? 1 : -1; }
    SP_SYNC();
    return s_int[2];
  }
  // the sequential walk needs the two special cases first: they run the parallel steps again on sub-strings
  int nout = 0;
  int i0 = 0, pred_r = 0, pred_a = 0, pred_b = 0;
  {
    const int f = sc.bl[0];
    const int sr = sc.ar[f] + k, sa = sc.aa[f] + k, sb = sc.ab[f] + k;   // start_ref / start_S1 / start_S2 lengths (:264-266; substr clamps)
    const int lr = sr < ref.n ? sr : ref.n, la = sa < S1.n ? sa : S1.n, lb = sb < S2.n ? sb : S2.n;
    if ((long long)lb * 2 < lr && (unsigned)(lr - lb) > 200u) {
      // the corrected read starts late: reference and uncorrected prefix are cut on their own (S2 := the reference prefix),
      // with a minimum window of 1.2 x the corrected prefix; the corrected side gets N records and its prefix last (:268-277)
      const SplitSeq pr{ref.s, lr}, pa{S1.s, la};
      const uint32_t ms = (uint32_t)(1.2 * (double)lb);
      const int bl2 = split_chain(sc2, pr, pa, pr, k, ms, s_int + 4);
      SP_SERIAL {
        int n2 = 0;
        if (bl2 < 1) { if (nout < out_cap) out[nout] = SplitWin{0, pr.n, 0, pa.n, 0, 0}; n2 = 1; }
        else {
          int qr = 0, qa = 0, qb = 0;
          for (int i = 0; i < bl2 - 1; ++i) {
            const int x = sc2.bl[i];
            const int zr = sc2.ar[x] - qr, za = sc2.aa[x] - qa, zb = sc2.ab[x] - qb;
            auto ab = [](int v) { return v < 0 ? -v : v; };
            if ((uint32_t)zr > ms && (uint32_t)za > ms && (uint32_t)zb > ms && (double)ab(za - zr) < zr * 0.5 && (double)ab(zb - zr) < zr * 0.5) {
              if (nout + n2 < out_cap) out[nout + n2] = SplitWin{qr, sc2.ar[x] - qr + k, qa, sc2.aa[x] - qa + k, 0, 0};
              ++n2;
              qr = sc2.ar[x] + k; qa = sc2.aa[x] + k; qb = sc2.ab[x] + k;
            }
          }
          if (nout + n2 < out_cap) out[nout + n2] = SplitWin{qr, pr.n - qr, qa, pa.n - qa, 0, 0};   // first_call is false there: the rest as it is (:302-306)
          ++n2;
        }
        // corrected side: n2 - 1 times N, then the prefix (N when it is empty) -- generate_dumb_str(n2, header, start_S2, "")
        for (int i = 0; i < n2 && nout + i < out_cap; ++i) { out[nout + i].b0 = -1; out[nout + i].bn = 1; }
        if (lb > 0 && nout + n2 - 1 < out_cap) { out[nout + n2 - 1].b0 = 0; out[nout + n2 - 1].bn = lb; }
        s_int[2] = n2;
      }
      SP_SYNC();
      nout += s_int[2];
      pred_r = sr; pred_a = sa; pred_b = sb;   // (:273-275: the anchor's end, unclamped)
      i0 = 1;
    }
  }
  // the walk (:280-290) and the tail (:292-306); the early-end case needs its own parallel steps, so the walk stops before it
  SP_SERIAL {
    int n = nout;
    for (int i = i0; i < blen - 1; ++i) {
      const int x = sc.bl[i];
      const int zr = sc.ar[x] - pred_r, za = sc.aa[x] - pred_a, zb = sc.ab[x] - pred_b;
      auto ab = [](int v) { return v < 0 ? -v : v; };
      if ((uint32_t)zr > min_size && (uint32_t)za > min_size && (uint32_t)zb > min_size && (double)ab(za - zr) < zr * 0.5 && (double)ab(zb - zr) < zr * 0.5) {
        if (n < out_cap) out[n] = SplitWin{pred_r, zr + k, pred_a, za + k, pred_b, zb + k};
        ++n;
        pred_r = sc.ar[x] + k; pred_a = sc.aa[x] + k; pred_b = sc.ab[x] + k;
      }
    }
    s_int[2] = n; s_int[3] = pred_r; s_int[8] = pred_a; s_int[9] = pred_b;
  }
  SP_SYNC();
  nout = s_int[2]; pred_r = s_int[3]; pred_a = s_int[8]; pred_b = s_int[9];
  {
    // substr(pred) of a read that is shorter than pred cannot happen on the chain (positions + k <= length)
    const int er = ref.n - pred_r, ea = S1.n - pred_a, eb = S2.n - pred_b;
    if ((long long)eb * 2 < er && (unsigned)(er - eb) > 200u) {
      // the corrected read ends early (:295-301): the corrected side gets its tail first, then N records
      const SplitSeq pr{ref.s + pred_r, er}, pa{S1.s + pred_a, ea};
      const uint32_t ms = (uint32_t)(1.2 * (double)eb);
      const int bl2 = split_chain(sc2, pr, pa, pr, k, ms, s_int + 4);
      SP_SERIAL {
        int n2 = 0;
        if (bl2 < 1) { if (nout < out_cap) out[nout] = SplitWin{pred_r, er, pred_a, ea, 0, 0}; n2 = 1; }
        else {
          int qr = 0, qa = 0, qb = 0;
          for (int i = 0; i < bl2 - 1; ++i) {
            const int x = sc2.bl[i];
            const int zr = sc2.ar[x] - qr, za = sc2.aa[x] - qa, zb = sc2.ab[x] - qb;
            auto ab = [](int v) { return v < 0 ? -v : v; };
            if ((uint32_t)zr > ms && (uint32_t)za > ms && (uint32_t)zb > ms && (double)ab(za - zr) < zr * 0.5 && (double)ab(zb - zr) < zr * 0.5) {
              if (nout + n2 < out_cap) out[nout + n2] = SplitWin{pred_r + qr, sc2.ar[x] - qr + k, pred_a + qa, sc2.aa[x] - qa + k, 0, 0};
              ++n2;
              qr = sc2.ar[x] + k; qa = sc2.aa[x] + k; qb = sc2.ab[x] + k;
            }
          }
          if (nout + n2 < out_cap) out[nout + n2] = SplitWin{pred_r + qr, er - qr, pred_a + qa, ea - qa, 0, 0};
          ++n2;
        }
        for (int i = 0; i < n2 && nout + i < out_cap; ++i) { out[nout + i].b0 = -1; out[nout + i].bn = 1; }
        if (eb > 0 && nout < out_cap) { out[nout].b0 = pred_b; out[nout].bn = eb; }
        s_int[2] = nout + n2;
      }
    } else {
      SP_SERIAL {
        if (nout < out_cap) out[nout] = SplitWin{pred_r, er, pred_a, ea, pred_b, eb};
        s_int[2] = nout + 1;
      }
    }
  }
  SP_SYNC();
  nout = s_int[2];
  return nout <= out_cap ? nout : -1;
}

}  // namespace elector

#ifdef __CUDACC__
namespace elector {

struct SplitArgs {
  int64_t n_triplets;
  const uint8_t *let[3];        // ref, S1 (uncorrected), S2 (corrected) letters of the call
  const int64_t *off[3];        // n_triplets + 1 each
  const int32_t *header_len;
  const int32_t *status_in;     // 1: corrected read shorter than the threshold share of the reference (no job)
  // scratch: per CTA two tables + anchor arrays, sized for the longest reference read of the call
  uint32_t *pool; uint64_t cta_words; uint32_t max_slots; int32_t max_anchors; uint32_t cand_words;
  // per (triplet, k) job: records, their number (-1: capacity), largest_fragment()
  SplitWin *wins; const int64_t *win_off;   // job (t, ki) at wins[win_off[t] + ki * cap(t)], cap(t) = (win_off[t+1] - win_off[t]) / 4
  int32_t *job_n; uint32_t *job_largest;
  int32_t *counter;
};

__device__ inline SplitScratch carve_scratch(uint32_t *base, uint32_t max_slots, uint32_t cand_words, int32_t max_anchors, uint32_t slots) {
  SplitScratch sc;
  uint32_t *p = base;
  sc.key = p; p += max_slots; sc.flag = p; p += max_slots; sc.posr = p; p += max_slots; sc.posa = p; p += max_slots; sc.posb = p; p += max_slots;
  sc.cand = p; p += cand_words;
  sc.ar = reinterpret_cast<int32_t *>(p); p += max_anchors; sc.aa = reinterpret_cast<int32_t *>(p); p += max_anchors;
  sc.ab = reinterpret_cast<int32_t *>(p); p += max_anchors; sc.chain = reinterpret_cast<int32_t *>(p); p += max_anchors;
  sc.nxt = reinterpret_cast<int32_t *>(p); p += max_anchors; sc.bl = reinterpret_cast<int32_t *>(p);
  sc.slots = slots; sc.max_anchors = max_anchors;
  return sc;
}
inline uint64_t split_scratch_words(uint32_t max_slots, uint32_t cand_words, int32_t max_anchors) { return 5ull * max_slots + cand_words + 6ull * (uint64_t)max_anchors; }

// persistent CTAs; a job = (triplet, k); the table of a job has 2 slots per k-mer of ITS reference read
__global__ void __launch_bounds__(256) split_jobs_kernel(SplitArgs a) {
  __shared__ int s_int[16];
  __shared__ int s_job;
  uint32_t *base = a.pool + (uint64_t)blockIdx.x * a.cta_words;
  const uint64_t half = a.cta_words / 2;
  for (;;) {
    if (threadIdx.x == 0) s_job = atomicAdd(a.counter, 1);
    __syncthreads();
    const int64_t job = s_job;
    __syncthreads();
    if (job >= 4 * a.n_triplets) break;
    const int64_t t = job >> 2;
    const int ki = (int)(job & 3), k = 15 - 2 * ki;
    if (a.status_in[t]) { if (threadIdx.x == 0) { a.job_n[job] = 0; a.job_largest[job] = 0; } continue; }
    const SplitSeq ref{a.let[0] + a.off[0][t], (int)(a.off[0][t + 1] - a.off[0][t])}, S1{a.let[1] + a.off[1][t], (int)(a.off[1][t + 1] - a.off[1][t])},
        S2{a.let[2] + a.off[2][t], (int)(a.off[2][t + 1] - a.off[2][t])};
    uint32_t slots = 64;
    while (slots < 2u * (uint32_t)ref.n + 2u) slots <<= 1;
    if (slots > a.max_slots) slots = a.max_slots;
    const SplitScratch sc = carve_scratch(base, a.max_slots, a.cand_words, a.max_anchors, slots);
    const SplitScratch sc2 = carve_scratch(base + half, a.max_slots, a.cand_words, a.max_anchors, slots);
    const int64_t cap = (a.win_off[t + 1] - a.win_off[t]) >> 2;
    SplitWin *out = a.wins + a.win_off[t] + (int64_t)ki * cap;
    const int n = split_job(sc, sc2, ref, S1, S2, k, out, (int)cap, s_int);
    if (threadIdx.x == 0) {
      a.job_n[job] = n;
      unsigned largest = (unsigned)a.header_len[t] + (n > 1 ? 1u : 0u);     // largest_fragment(), :158-169 (split_host.hpp)
      for (int i = 0; i < n; ++i) largest = max(largest, (unsigned)out[i].rn + 1u);
      a.job_largest[job] = largest;
    }
    __syncthreads();
  }
}

// best_split (:310-332): k = 15, then 13, 11, 9 as long as the largest fragment gets strictly smaller.  One thread per triplet:
// the chosen k, the status (2: at most one record -> the AAA placeholder, :417-422), the records and letters of each kind it
// contributes.
__global__ void split_select_kernel(int64_t n, const int32_t *status_in, const int32_t *job_n, const uint32_t *job_largest, const SplitWin *wins,
                                    const int64_t *win_off, int32_t *status, int32_t *k_idx, int64_t *nrec, int64_t *nlet_r, int64_t *nlet_a, int64_t *nlet_b,
                                    int32_t *error) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int st = status_in[t], best = 0;
  if (!st) {
    for (int ki = 0; ki < 4; ++ki) if (job_n[4 * t + ki] < 0) atomicExch(error, 1);
    for (int ki = 1; ki < 4; ++ki) { if (job_largest[4 * t + ki] < job_largest[4 * t + best]) best = ki; else break; }
    if (job_n[4 * t + best] <= 1) st = 2;
  }
  status[t] = st; k_idx[t] = best;
  int64_t lr = 3, la = 3, lb = 3, nr = 1;   // the placeholder record AAA / AAA / AAA
  if (!st) {
    const int64_t cap = (win_off[t + 1] - win_off[t]) >> 2;
    const SplitWin *w = wins + win_off[t] + (int64_t)best * cap;
    nr = job_n[4 * t + best]; lr = la = lb = 0;
    for (int i = 0; i < nr; ++i) { lr += w[i].rn; la += w[i].an; lb += w[i].bn; }
  }
  nrec[t] = nr; nlet_r[t] = lr; nlet_a[t] = la; nlet_b[t] = lb;
}

// the windows of the call as three letter arrays with offsets (what `poa` reads from the shard files): one CTA per triplet
__global__ void __launch_bounds__(128) split_emit_kernel(int64_t n, const uint8_t *l0, const uint8_t *l1, const uint8_t *l2, const int64_t *o0, const int64_t *o1,
                                                          const int64_t *o2, const int32_t *status, const int32_t *k_idx, const SplitWin *wins, const int64_t *win_off,
                                                          const int64_t *rec_base, const int64_t *br, const int64_t *ba, const int64_t *bb, int64_t *w_off_r,
                                                          int64_t *w_off_a, int64_t *w_off_b, uint8_t *w_r, uint8_t *w_a, uint8_t *w_b, int64_t *read_first) {
  const int64_t t = blockIdx.x;
  if (t >= n) return;
  const int64_t w0 = rec_base[t];
  if (threadIdx.x == 0) { read_first[t] = w0; if (t == n - 1) read_first[n] = rec_base[n]; }
  if (status[t]) {
    if (threadIdx.x < 3) { w_r[br[t] + threadIdx.x] = 'A'; w_a[ba[t] + threadIdx.x] = 'A'; w_b[bb[t] + threadIdx.x] = 'A'; }
    if (threadIdx.x == 0) { w_off_r[w0] = br[t]; w_off_a[w0] = ba[t]; w_off_b[w0] = bb[t]; if (t == n - 1) { w_off_r[w0 + 1] = br[t] + 3; w_off_a[w0 + 1] = ba[t] + 3; w_off_b[w0 + 1] = bb[t] + 3; } }
    return;
  }
  const int64_t cap = (win_off[t + 1] - win_off[t]) >> 2;
  const SplitWin *w = wins + win_off[t] + (int64_t)k_idx[t] * cap;
  const int nr = (int)(rec_base[t + 1] - w0);
  __shared__ int64_t s_cur[3];
  if (threadIdx.x == 0) {   // offsets of the records: a running sum over at most a few hundred of them
    int64_t cr = br[t], ca = ba[t], cb = bb[t];
    for (int i = 0; i < nr; ++i) { w_off_r[w0 + i] = cr; w_off_a[w0 + i] = ca; w_off_b[w0 + i] = cb; cr += w[i].rn; ca += w[i].an; cb += w[i].bn; }
    if (t == n - 1) { w_off_r[w0 + nr] = cr; w_off_a[w0 + nr] = ca; w_off_b[w0 + nr] = cb; }
    s_cur[0] = br[t]; s_cur[1] = ba[t]; s_cur[2] = bb[t];
  }
  __syncthreads();
  // reference and uncorrected windows tile their reads: the letters of the triplet are one contiguous copy each (the records of
  // the two special cases are in read order too); the corrected side has N records in between
  const uint8_t *sr = l0 + o0[t], *sa = l1 + o1[t];
  {
    int64_t tot_r = 0, tot_a = 0;
    for (int i = 0; i < nr; ++i) { tot_r += w[i].rn; tot_a += w[i].an; }
    const int first_r = w[0].r0, first_a = w[0].a0;
    for (int64_t i = threadIdx.x; i < tot_r; i += blockDim.x) w_r[s_cur[0] + i] = sr[first_r + i];
    for (int64_t i = threadIdx.x; i < tot_a; i += blockDim.x) w_a[s_cur[1] + i] = sa[first_a + i];
  }
  const uint8_t *sb = l2 + o2[t];
  for (int i = 0; i < nr; ++i) {
    const int64_t dst = w_off_b[w0 + i];
    if (w[i].b0 < 0) { if (threadIdx.x == 0) w_b[dst] = 'N'; }
    else for (int j = threadIdx.x; j < w[i].bn; j += blockDim.x) w_b[dst + j] = sb[w[i].b0 + j];
  }
}

}  // namespace elector
#endif

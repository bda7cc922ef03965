// split_kernel.cuh -- ELECTOR's window cutting (src/split/Master_Splitter.cpp) on the device: SURVEY.md 8(f)-1.
//
// The reference cuts every (reference, uncorrected, corrected) read triplet into ~50-letter windows at k-mers that occur
// exactly once in each of the three reads (split() :175-308), keeps the cutting with the smallest largest window over
// k = 15, 13, 11, 9 (best_split() :310-332), and hands the windows to `poa`.  It is a serial 10.7 ms per triplet; with the
// alignment at 0.7 us per triplet it is 99.9 % of the stage.
//
// Mapping: ONE CTA PER (triplet, k) JOB -- the four k of a triplet run side by side, the choice between them is made
// afterwards.  A job lives in the CTA's shared memory (reads of up to ~14 000 letters with two CTAs per SM, ~28 000 with one;
// longer ones take the same code through a pool in global memory).  Per job:
//   0. the three reads are packed to 2 bits per letter (a k-mer is then two word loads and a funnel shift);
//   1. the k-mers of the reference go into an open-addressing hash table of 4-byte entries: the position of the k-mer's first
//      occurrence in the reference (the key is read back from the packed read) and six flags "seen once" / "seen again" per
//      read, set with atomicOr -- "exactly once in each read" is a flag pattern (:176-230 keep three std::unordered_maps
//      with -1 for repeats);
//   2. the k-mers of the other two reads look their entry up and set their flags;
//   3. a sweep over the table sets the bit of every such k-mer's reference position -> the candidates in reference order;
//   4. one thread thins the bitmap like the reference's left-to-right scan (:242-251, an anchor at most every minSize + 1
//      letters) by jumping from anchor to the next set bit; the chosen anchors' entries are marked and a second pass over the
//      other two reads picks their positions up (no position arrays beside the table);
//   5. the longest chain of anchors that increase by less than 1000 letters in all three reads (:79-126, a memoised
//      recursion) is a backward dynamic programme over the anchor list: a warp looks at the <= 48 successors of an anchor
//      at once (one redux.sync for "first maximum", like the reference's strict comparisons);
//   6. one thread walks the chain and cuts (:262-306), including the two special cases of a corrected read that starts late
//      or ends early (the reference then splits reference and uncorrected read alone and pads the corrected side with `N`
//      records, :268-277 and :295-301): the CTA runs steps 0-5 again on the sub-strings.
// The same code compiles for the host (tests/emul/split_emul.cu runs it serially against the compiled reference).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define SP_HD __host__ __device__ __forceinline__   // inlined into the kernel: the compiler then sees which arrays are shared memory
#else
#define SP_HD inline
#endif

namespace elector {

constexpr uint32_t kSplitEmpty = 0xffffffffu;
enum : uint32_t { SF_R1 = 1, SF_R2 = 2, SF_A1 = 4, SF_A2 = 8, SF_B1 = 16, SF_B2 = 32, SF_MASK = 63, SF_ONCE = SF_R1 | SF_A1 | SF_B1 };   // seen once / again in ref, S1, S2
constexpr int kChainReach = 1000;              // Master_Splitter.cpp:86-88
constexpr int kSplitMaxRead = (1 << 25) - 2;   // a table entry is fingerprint | reference position | 6 flags

// a read (or a piece of one): its letters, and -- on the device -- where they sit in the call's letters packed once at 2 bits each
// (split_prepack_kernel): letter g + t of the kind's array is letter t of this read
struct SplitSeq { const uint8_t *s; int n; const uint32_t *pk = nullptr; int64_t g = 0; };

// one output record of a job: three (start, length) pairs into the job's three reads; b0 = -1: the corrected side is the
// placeholder "N" (generate_dumb_str, :139-154)
struct SplitWin { int32_t r0, rn, a0, an, b0, bn; };

// an anchor: positions of the k-mer in the three reads, and the length of the best chain that starts at it
struct alignas(16) SplitAnchor { int32_t r, a, b, chain; };

// scratch of one job: shared memory when the job fits the CTA's share (every array then), the CTA's pool in global memory when
// it does not (reads of more than ~13 000 / ~27 000 letters)
struct SplitScratch {
  uint32_t *table;      // open addressing in buckets of 4 slots (one 16-byte load per probe), `slots` entries: kSplitEmpty or
  uint32_t slots;       // fingerprint | position of the k-mer's first occurrence in ref | flags; a probe compares the fingerprints and only
                        // then the k-mer itself, read from the packed reference.  A bucket fills from its first slot on: an empty last
                        // slot ends a search.  After step 3 the table is dead and its words hold the thinning's pointers, then the chain
                        // DP's eligibility masks
  uint32_t pos_bits;    // width of the position field above the 6 flag bits
  uint32_t *pk[3];      // the three reads at 2 bits per letter, 16 letters per word (+ 2 words that a k-mer at the end may touch)
  uint32_t *cand;       // bitmap over reference positions
  SplitAnchor *anc;     // anchors in reference order
  int32_t *nxt;         // successor on the best chain
  int32_t *bl;          // the chain: indices into the anchor list
  int32_t max_anchors;
  uint32_t *t2_km;      // a second, small table of the chosen anchors' k-mers (t2_slots entries) ...
  int32_t *t2_idx;      // ... and their indices: the other two reads are looked up in it for the anchors' positions
  uint32_t t2_slots;
};

// words of scratch that a job needs beside its table
SP_HD uint32_t split_pk_words(int n) { return (uint32_t)(n >> 4) + 2u; }
SP_HD uint32_t split_cand_words(int n) { return (uint32_t)(n >> 5) + 2u; }
SP_HD int32_t split_anchor_bound(int n, uint32_t min_size) { return (int32_t)((((uint32_t)n / (min_size + 1u)) + 6u) & ~3u); }   // anchors are more than min_size apart (:242-251)
SP_HD uint32_t split_min_slots(int n) { return (uint32_t)n + (uint32_t)(n >> 2) + (uint32_t)(n >> 4) + 64u; }                     // load factor <= 0.76
SP_HD uint64_t split_fixed_words(int nr, int na, int nb, int32_t max_anchors, int32_t sub_anchors) {
  return (uint64_t)split_pk_words(nr) + split_pk_words(na) + split_pk_words(nb) + split_cand_words(nr) + 6ull * (uint64_t)max_anchors + 6ull * (uint64_t)sub_anchors +
         4ull * (uint64_t)(max_anchors > sub_anchors ? max_anchors : sub_anchors);
}
// Places the arrays of a job in `base` (`words` 32-bit words) -- sc for the job's own chain, sc2 for the sub-calls of the two
// special cases (same table and packed reads: the outer table is no longer needed then; own anchor arrays: the outer anchors
// are) -- and says whether they fit.  The table gets what is left, at most 4 slots per reference k-mer.  na / nb: the longest
// second / third read the scratch will see (a sub-call puts the reference prefix in the third place).  max_anchors and
// sub_anchors are multiples of 4.
SP_HD bool split_carve(SplitScratch &sc, SplitScratch &sc2, uint32_t *base, uint64_t words, int nr, int na, int nb, int32_t max_anchors, int32_t sub_anchors) {
  const uint64_t fixed = split_fixed_words(nr, na, nb, max_anchors, sub_anchors);
  if (words < fixed + split_min_slots(nr)) return false;
  const int32_t most_anchors = max_anchors > sub_anchors ? max_anchors : sub_anchors;
  uint32_t *p = base;
  sc.anc = reinterpret_cast<SplitAnchor *>(p); p += 4 * (size_t)max_anchors;     // 16-byte aligned: first
  SplitAnchor *anc2 = reinterpret_cast<SplitAnchor *>(p); p += 4 * (size_t)sub_anchors;
  sc.nxt = reinterpret_cast<int32_t *>(p); p += max_anchors;
  sc.bl = reinterpret_cast<int32_t *>(p); p += max_anchors;
  int32_t *nxt2 = reinterpret_cast<int32_t *>(p); p += sub_anchors;
  int32_t *bl2 = reinterpret_cast<int32_t *>(p); p += sub_anchors;
  sc.t2_km = p; p += 2 * (size_t)most_anchors; sc.t2_idx = reinterpret_cast<int32_t *>(p); p += 2 * (size_t)most_anchors;
  sc.t2_slots = 2u * (uint32_t)most_anchors;
  sc.table = p;                                                                   // 16-byte aligned: everything before it is a multiple of 4 words
  const uint64_t left = words - fixed, most = 4ull * (uint64_t)nr + 64ull;
  sc.slots = (uint32_t)(left < most ? left : most) & ~3u;                          // buckets of 4 slots
  p += sc.slots;
  sc.pk[0] = p; p += split_pk_words(nr); sc.pk[1] = p; p += split_pk_words(na); sc.pk[2] = p; p += split_pk_words(nb);
  sc.cand = p; p += split_cand_words(nr);
  sc.pos_bits = 1;
  while ((1u << sc.pos_bits) <= (uint32_t)nr) ++sc.pos_bits;                      // 2^pos_bits > nr: a position field is never all ones
  sc.max_anchors = max_anchors;
  sc2 = sc;
  sc2.anc = anc2; sc2.nxt = nxt2; sc2.bl = bl2; sc2.max_anchors = sub_anchors;
  return true;
}

// letter codes of the reference's two encoders: str2num (:20-32) for the first k letters of a read, nuc2int (:35-43) for the rest
SP_HD uint32_t split_code(const SplitSeq &q, int t, int k) {
  const uint8_t c = q.s[t];
  if (t < k) return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;
  return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}
SP_HD int split_kmers(const SplitSeq &q, int k) { return q.n > k ? q.n - k + 1 : 1; }
// The k-mer at position p of a packed read: letter p in the lowest two bits.  The reference keeps the first letter in the highest
// bits (:22-30,:54-57); k-mers are only ever compared, so any one-to-one code gives the same anchors.  A read of at most k letters
// has one shorter "k-mer" (substr(0, k), :177) whose value is that of the k-mer with leading A's: the letters sit at the top here.
SP_HD uint32_t split_kmer(const uint32_t *pk, int p, int k, int n) {
  const uint32_t lo = pk[p >> 4], hi = pk[(p >> 4) + 1];
  const int sh = 2 * (p & 15);
#ifdef __CUDA_ARCH__
  const uint32_t v = __funnelshift_r(lo, hi, sh);
#else
  const uint32_t v = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
  if (n >= k) return v & ((1u << (2 * k)) - 1u);
  return (v & ((1u << (2 * n)) - 1u)) << (2 * (k - n));
}
// the same for a read of at least k letters, with the mask (1 << 2k) - 1 at hand
SP_HD uint32_t split_kmer_full(const uint32_t *pk, int p, uint32_t mask) {
  const uint32_t lo = pk[p >> 4], hi = pk[(p >> 4) + 1];
  const int sh = 2 * (p & 15);
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh) & mask;
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & mask;
#endif
}
SP_HD uint32_t split_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
SP_HD uint32_t split_mix(uint32_t kmer) { uint32_t h = kmer * 0x9E3779B1u; h ^= h >> 15; return h * 0x85EBCA77u; }
// the fingerprint field of an entry for hash h (fsh = 6 + pos_bits): never all ones, so that kSplitEmpty fails every fingerprint test
SP_HD uint32_t split_want(uint32_t h, uint32_t fsh) { const uint32_t w = (h >> 3) << fsh; return w == (0xffffffffu << fsh) ? 0u : w; }

#if defined(SPLIT_TIMING) && defined(__CUDACC__)
// phase clocks of the cutting kernel (a diagnostic build: -DSPLIT_TIMING): cycles of thread 0 between the marks, summed over jobs
__device__ unsigned long long g_split_clk[16];
__host__ __device__ __forceinline__ void split_mark(int k, long long &prev) {
#ifdef __CUDA_ARCH__
  if (threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&g_split_clk[k], (unsigned long long)(t - prev)); prev = t; }
#endif
}
#define ST(k) split_mark(k, *reinterpret_cast<long long *>(s_int + 12))
#else
#define ST(k)
#endif

#ifdef __CUDA_ARCH__
#define SP_FIRST ((int)threadIdx.x)
#define SP_STRIDE ((int)blockDim.x)
#define SP_FOR(i, n) for (int i = threadIdx.x; i < (int)(n); i += blockDim.x)
#define SP_SYNC() __syncthreads()
#define SP_SERIAL if (threadIdx.x == 0)
SP_HD uint32_t sp_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
SP_HD uint32_t sp_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
#else
#define SP_FIRST 0
#define SP_STRIDE 1
#define SP_FOR(i, n) for (int i = 0; i < (int)(n); ++i)
#define SP_SYNC()
#define SP_SERIAL
SP_HD uint32_t sp_cas(uint32_t *p, uint32_t cmp, uint32_t v) { const uint32_t o = *p; if (o == cmp) *p = v; return o; }
SP_HD uint32_t sp_or(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o | v; return o; }
#endif

// 16 letters of q from position 16 * w as one word: from the bytes, or from the packed letters of the call (every letter coded
// like nuc2int there; the first k letters of a read take str2num's code, so the first word is patched from the bytes)
SP_HD uint32_t split_pack_word(const SplitSeq &q, int w, int k) {
  uint32_t v = 0;
  const int t0 = 16 * w, e = t0 + 16 < q.n ? t0 + 16 : q.n;
  if (q.pk) {
    if (e <= t0) return 0;
    const int64_t g = q.g + t0;
    const uint32_t lo = q.pk[g >> 4], hi = q.pk[(g >> 4) + 1];
    const int sh = 2 * (int)(g & 15);
#ifdef __CUDA_ARCH__
    v = __funnelshift_r(lo, hi, sh);
#else
    v = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
    if (e - t0 < 16) v &= (1u << (2 * (e - t0))) - 1u;
    if (w == 0) { const int f = k < e ? k : e; for (int t = 0; t < f; ++t) v = (v & ~(3u << (2 * t))) | (split_code(q, t, k) << (2 * t)); }
    return v;
  }
  for (int t = t0; t < e; ++t) v |= split_code(q, t, k) << (2 * (t - t0));
  return v;
}

struct alignas(16) SplitBucket { uint32_t e[4]; };

// may the record that ends at an anchor zr / za / zb letters after the last cut be written (:282-283)?  The reference compares
// doubles (|za - zr| < zr * 0.5); the values are small integers, so twice the difference against zr is the same test.
SP_HD bool split_cut_ok(int zr, int za, int zb, uint32_t ms) {
  const int da = za > zr ? za - zr : zr - za, db = zb > zr ? zb - zr : zr - zb;     // positions are below 2^25: no overflow
  return (uint32_t)zr > ms && (uint32_t)za > ms && (uint32_t)zb > ms && 2 * da < zr && 2 * db < zr;
}

// Steps 0-5 for (ref, S1, S2): the anchor list and its best chain.  Returns the chain length (0: no anchor at all, -1: more anchors
// than the scratch holds) in every thread; the chain is sc.bl[0 .. len-1].  s_int: four ints of shared scratch (device) / any four
// ints (host).
SP_HD int split_chain(const SplitScratch &sc, const SplitSeq &ref, const SplitSeq &S1, const SplitSeq &S2, int k, uint32_t min_size, int *s_int) {
  const int nr = split_kmers(ref, k), na = split_kmers(S1, k), nb = split_kmers(S2, k);
  // 0. the reads at 2 bits per letter, empty tables
  SP_FOR(w, split_pk_words(ref.n)) sc.pk[0][w] = split_pack_word(ref, w, k);
  SP_FOR(w, split_pk_words(S1.n)) sc.pk[1][w] = split_pack_word(S1, w, k);
  SP_FOR(w, split_pk_words(S2.n)) sc.pk[2][w] = split_pack_word(S2, w, k);
  SP_FOR(i, sc.slots) sc.table[i] = kSplitEmpty;
  SP_FOR(i, sc.t2_slots) sc.t2_km[i] = kSplitEmpty;
  SP_FOR(i, split_cand_words(ref.n)) sc.cand[i] = 0;
  SP_SYNC(); ST(0);
  const uint32_t fsh = 6u + sc.pos_bits, pmask = (1u << sc.pos_bits) - 1u, flim = 1u << fsh, kmask = (1u << (2 * k)) - 1u;
  // k-mers of the three reads: two loads, a funnel shift and a mask when the read has at least k letters
  const bool full0 = ref.n >= k, full1 = S1.n >= k, full2 = S2.n >= k;
  auto vkmer = [&](const uint32_t *pk, int p) { return full0 ? split_kmer_full(pk, p, kmask) : split_kmer(pk, p, k, ref.n); };
  // 1. table of the reference k-mers (:176-193: an unordered_map with -1 for repeats)
  // (steps 1 and 2 are flat loops: a lane goes on to its next k-mer as soon as its own probe sequence ends, it does not wait for
  // the longest sequence of its warp)
  const uint32_t nbuckets = sc.slots >> 2;
  {
    int p = SP_FIRST;
    bool fresh = true;
    uint32_t km = 0, want = 0, mine = 0, b = 0;
    while (p < nr) {
      if (fresh) {
        km = vkmer(sc.pk[0], p);
        const uint32_t h = split_mix(km);
        want = split_want(h, fsh); mine = want | ((uint32_t)p << 6) | SF_R1; b = split_mulhi(h, nbuckets);
        fresh = false;
      }
      const SplitBucket v = reinterpret_cast<const SplitBucket *>(sc.table)[b];
      bool done = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (done) continue;
        uint32_t cur = v.e[j];
        if (cur == kSplitEmpty) {
          cur = sp_cas(&sc.table[4u * b + (uint32_t)j], kSplitEmpty, mine);
          if (cur == kSplitEmpty) { done = true; continue; }
        }
        if ((cur ^ want) < flim && vkmer(sc.pk[0], (int)((cur >> 6) & pmask)) == km) {
          if (!(cur & SF_R2)) sp_or(&sc.table[4u * b + (uint32_t)j], SF_R2);
          done = true;
        }
      }
      if (done) { p += SP_STRIDE; fresh = true; }
      else if (++b == nbuckets) b = 0;
    }
  }
  SP_SYNC(); ST(1);
  // 2. the other two reads (:194-230), one index space: 0 .. na - 1 is S1, na .. na + nb - 1 is S2
  {
    const int total = na + nb;
    int p = SP_FIRST;
    bool fresh = true;
    uint32_t km = 0, want = 0, fl = 0, b = 0;
    while (p < total) {
      if (fresh) {
        const bool first = p < na;
        km = first ? (full1 ? split_kmer_full(sc.pk[1], p, kmask) : split_kmer(sc.pk[1], p, k, S1.n))
                   : (full2 ? split_kmer_full(sc.pk[2], p - na, kmask) : split_kmer(sc.pk[2], p - na, k, S2.n));
        const uint32_t h = split_mix(km);
        want = split_want(h, fsh); fl = first ? (uint32_t)SF_A1 : (uint32_t)SF_B1; b = split_mulhi(h, nbuckets);
        fresh = false;
      }
      const SplitBucket v = reinterpret_cast<const SplitBucket *>(sc.table)[b];
      int hit = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (hit < 0 && (v.e[j] ^ want) < flim && vkmer(sc.pk[0], (int)((v.e[j] >> 6) & pmask)) == km) hit = j;
      if (hit >= 0) {
        const uint32_t s = 4u * b + (uint32_t)hit;
        if (sp_or(&sc.table[s], fl) & fl) sp_or(&sc.table[s], fl << 1);     // SF_A2 / SF_B2: seen again
        p += SP_STRIDE; fresh = true;
      } else if (v.e[3] == kSplitEmpty) { p += SP_STRIDE; fresh = true; }
      else if (++b == nbuckets) b = 0;
    }
  }
  SP_SYNC(); ST(2);
  // 3. anchors in reference order: once in each read
  SP_FOR(s, sc.slots) {
    const uint32_t e = sc.table[s];
    if (e != kSplitEmpty && (e & SF_MASK) == SF_ONCE) { const uint32_t r = (e >> 6) & pmask; sp_or(&sc.cand[r >> 5], 1u << (r & 31)); }
  }
  SP_SYNC(); ST(3);
  // 4. left-to-right thinning (:242-251): position 0 is taken as it is; position j + 1 when j - last > min_size, last = j -- the
  //    next anchor is the first candidate more than min_size after the last one (min_size + 1 after the start).  Every candidate
  //    learns its successor in parallel (in the words of the table, which is dead now); one thread then only follows pointers.
  const uint32_t cwords = (uint32_t)(nr + 31) >> 5;
  int32_t *jump = reinterpret_cast<int32_t *>(sc.table);
  auto first_from = [&](uint64_t pos) -> int32_t {             // first candidate at or after pos, -1 when there is none
    if (pos >= (uint64_t)nr) return -1;
    uint32_t w = (uint32_t)(pos >> 5);
    uint32_t m = sc.cand[w] & (0xffffffffu << (pos & 31));
    while (!m && ++w < cwords) m = sc.cand[w];
    if (!m) return -1;
    int b = 0;
#ifdef __CUDA_ARCH__
    b = __ffs((int)m) - 1;
#else
    while (!((m >> b) & 1u)) ++b;
#endif
    return (int32_t)(w * 32u) + b;
  };
  SP_FOR(w, cwords) {
    uint32_t bits = sc.cand[w];
    while (bits) {
      int b = 0;
#ifdef __CUDA_ARCH__
      b = __ffs((int)bits) - 1;
#else
      while (!((bits >> b) & 1u)) ++b;
#endif
      bits &= bits - 1;
      const int c = w * 32 + b;
      jump[c] = first_from((uint64_t)c + min_size + 1u);
    }
  }
  SP_SYNC(); ST(4);
  SP_SERIAL {
    int n = 0;
    bool over = false;
    if (sc.cand[0] & 1u) sc.anc[n++].r = 0;
    int32_t p = first_from((uint64_t)min_size + 2u);
    while (p >= 0) {
      if (n >= sc.max_anchors) { over = true; break; }
      sc.anc[n++].r = p;
      p = jump[p];
    }
    s_int[0] = over ? -1 : n;
  }
  SP_SYNC(); ST(5);
  const int n = s_int[0];
  if (n < 0) return -1;
  // the anchors' k-mers go into the small table; the other two reads are looked up in it once more, for the positions.  A one-hash
  // Bloom bitmap (the words of the candidate bitmap, free now) answers most of those look-ups with one load.
  const uint32_t bloom_bits = split_cand_words(ref.n) * 32u;
  SP_FOR(i, split_cand_words(ref.n)) sc.cand[i] = 0;
  SP_SYNC(); ST(6);
  SP_FOR(i, n) {
    const uint32_t km = split_kmer(sc.pk[0], sc.anc[i].r, k, ref.n), h = km * 0x9E3779B1u;
    uint32_t s = split_mulhi(h, sc.t2_slots);
    while (sp_cas(&sc.t2_km[s], kSplitEmpty, km) != kSplitEmpty) { if (++s == sc.t2_slots) s = 0; }
    sc.t2_idx[s] = i;
    const uint32_t bit = split_mulhi(h ^ (h >> 13), bloom_bits);
    sp_or(&sc.cand[bit >> 5], 1u << (bit & 31));
    sc.anc[i].chain = 0;
  }
  SP_SYNC(); ST(7);
  for (int q = 1; q <= 2; ++q) {
    const int nq = q == 1 ? na : nb, lq = q == 1 ? S1.n : S2.n;
    SP_FOR(p, nq) {
      const uint32_t km = (q == 1 ? full1 : full2) ? split_kmer_full(sc.pk[q], p, kmask) : split_kmer(sc.pk[q], p, k, lq), h = km * 0x9E3779B1u;
      const uint32_t bit = split_mulhi(h ^ (h >> 13), bloom_bits);
      if (!((sc.cand[bit >> 5] >> (bit & 31)) & 1u)) continue;
      uint32_t s = split_mulhi(h, sc.t2_slots);
      for (;;) {
        const uint32_t e = sc.t2_km[s];
        if (e == kSplitEmpty) break;
        if (e == km) { if (q == 1) sc.anc[sc.t2_idx[s]].a = p; else sc.anc[sc.t2_idx[s]].b = p; break; }
        if (++s == sc.t2_slots) s = 0;
      }
    }
  }
  SP_SYNC(); ST(8);
  // 5. longest chain, backwards: chain[i] = 1 + max over the successors within reach (first maximum), 0 when there is none
#ifdef __CUDA_ARCH__
  const bool fast = (min_size + 1u) * 64u >= (uint32_t)kChainReach && 2u * (uint32_t)n <= sc.slots;
  uint32_t *elig = sc.table;                                   // bit j of anchor i's 64: anchor i + 1 + j may follow it (:86-88)
  if (fast) {
    const int total = n * 64, lane_ = threadIdx.x & 31;
    for (int t = threadIdx.x; t - lane_ < total; t += blockDim.x) {
      const int i = t >> 6, c = i + 1 + (t & 63);
      bool ok = false;
      if (t < total && c < n) {
        const SplitAnchor me = sc.anc[i], o = sc.anc[c];
        const int da = o.a - me.a, db = o.b - me.b;
        ok = o.r - me.r < kChainReach && da > 0 && da < kChainReach && db > 0 && db < kChainReach;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (lane_ == 0 && t < total) elig[t >> 5] = m;
    }
    __syncthreads();
  }
  if (fast) {
    // at most 63 successors are within reach and their eligibility (a pure function of the positions) is in the masks.  Blocks of
    // 32 anchors from the end, lane l = anchor b + l: first the whole CTA takes, for every anchor of the block, the maximum over its
    // successors BEHIND the block (their chain lengths are final), then one warp settles the block from its last anchor down -- one
    // shuffle per anchor hands its final length to the lanes before it.  key = (chain length of the successor + 1) << 6 |
    // (63 - distance): the maximum is the longest chain and, among equals, the nearest successor (the reference's first strict
    // maximum).
    const uint2 *elig2 = reinterpret_cast<const uint2 *>(elig);
    unsigned *skey = reinterpret_cast<unsigned *>(sc.t2_km);                // 32 words of the small table (free during the DP)
    if (threadIdx.x < 32) skey[threadIdx.x] = 0;
    __syncthreads();
    for (int b = ((n - 1) >> 5) << 5; b >= 0; b -= 32) {
      const int cend = b + 95 < n ? b + 95 : n, span = cend - (b + 32);
      for (int t = threadIdx.x; t < span * 32; t += blockDim.x) {           // a warp = one successor c, its lanes = the block's anchors
        const int c = b + 32 + (t >> 5), l = t & 31, i = b + l;
        const unsigned off = (unsigned)(c - i - 1);
        if (i < n && off < 63u) {
          const uint2 e = elig2[i];
          if (((off < 32u ? e.x >> off : e.y >> (off - 32u)) & 1u)) atomicMax(&skey[l], ((unsigned)(sc.anc[c].chain + 1) << 6) | (63u - off));
        }
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x, i = b + lane;
        const unsigned ex = i < n ? elig2[i].x : 0u;
        unsigned key = skey[lane];
        for (int t = 31; t >= 1; --t) {
          const unsigned vt = __shfl_sync(0xffffffffu, key >> 6, t) + 1u;    // final: every anchor behind b + t has been seen
          const unsigned off = (unsigned)(t - lane - 1);
          const unsigned cand = (vt << 6) | (63u - off);
          if (lane < t && ((ex >> off) & 1u) && cand > key) key = cand;
        }
        if (i < n) { sc.anc[i].chain = (int)(key >> 6); sc.nxt[i] = key ? i + 1 + 63 - (int)(key & 63u) : -1; }
        skey[lane] = 0;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (!fast) {
      for (int i = n - 1; i >= 0; --i) {
        const SplitAnchor me = sc.anc[i];
        int best = -1, arg = -1;
        for (int base = i + 1; base < n; base += 32) {          // successors in list order until the reference distance reaches 1000 (:86,:99)
          const int c = base + lane;
          unsigned key = 0;
          bool far = false;
          if (c < n) {
            const SplitAnchor o = sc.anc[c];
            far = !(o.r - me.r < kChainReach);
            const int da = o.a - me.a, db = o.b - me.b;
            if (!far && da > 0 && da < kChainReach && db > 0 && db < kChainReach) key = ((unsigned)(o.chain + 1) << 5) | (unsigned)(31 - lane);
          }
          // anchors are in reference order: once one is too far the rest is
          const unsigned top = __reduce_max_sync(0xffffffffu, key);
          if (top) { const int v = (int)(top >> 5) - 1; if (v > best) { best = v; arg = base + 31 - (int)(top & 31u); } }
          if (__any_sync(0xffffffffu, far)) break;
        }
        if (lane == 0) { sc.anc[i].chain = 1 + best; sc.nxt[i] = arg; }
        __syncwarp();
      }
    }
    // the first anchor with the longest chain starts it (:113-119)
    unsigned long long top = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const unsigned long long key = ((unsigned long long)(unsigned)(sc.anc[i].chain + 1) << 32) | (unsigned)(0x7fffffff - i);
      if (key > top) top = key;
    }
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, top, d); if (o > top) top = o; }
    if (lane == 0) s_int[1] = top ? 0x7fffffff - (int)(unsigned)(top & 0xffffffffu) : -1;
  }
  SP_SYNC(); ST(9);
  // the chain from there, by pointer doubling: round r places the anchors 2^r .. 2^(r+1) - 1 steps down the chain
  SP_FOR(i, n) sc.t2_idx[i] = sc.nxt[i];
  SP_SYNC(); ST(10);
  const int at = s_int[1];
  const int len = at < 0 ? 0 : sc.anc[at].chain + 1;
  if (len > 0) {
    int32_t *jc = sc.t2_idx, *jn = reinterpret_cast<int32_t *>(sc.t2_km);
    SP_SERIAL sc.bl[0] = at;
    SP_SYNC();
    for (int step = 1; step < len; step <<= 1) {
      const int fill = step < len - step ? step : len - step;
      SP_FOR(j, fill) sc.bl[j + step] = jc[sc.bl[j]];
      if (2 * step < len) SP_FOR(i, n) { const int t = jc[i]; jn[i] = t < 0 ? -1 : jc[t]; }
      SP_SYNC();
      int32_t *t = jc; jc = jn; jn = t;
    }
  }
  ST(11);
  return len;
#else
  for (int i = n - 1; i >= 0; --i) {
    int best = -1, arg = -1;
    for (int c = i + 1; c < n; ++c) {
      if (!(sc.anc[c].r - sc.anc[i].r < kChainReach)) break;
      const int da = sc.anc[c].a - sc.anc[i].a, db = sc.anc[c].b - sc.anc[i].b;
      if (da > 0 && da < kChainReach && db > 0 && db < kChainReach && sc.anc[c].chain > best) { best = sc.anc[c].chain; arg = c; }
    }
    sc.anc[i].chain = 1 + best; sc.nxt[i] = arg;
  }
  {   // the first anchor with the longest chain starts it (:113-119), then the successors
    int best = -1, at = -1;
    for (int i = 0; i < n; ++i) if (sc.anc[i].chain > best) { best = sc.anc[i].chain; at = i; }
    int len = 0;
    while (at != -1) { sc.bl[len++] = at; at = sc.nxt[at]; }
    s_int[1] = len;
  }
  return s_int[1];
#endif
}

// The walk over chain elements i0 .. i1 - 1 (:280-290): an anchor that may end a record (split_cut_ok) does, and the next record
// starts behind its k-mer.  q* : where the running record starts (in the walked strings), updated; off_*: where the walked strings
// start in the job's reads; with_b: the third read is cut too.  Records are written from out[*n] on, *n counts them all.
SP_HD void split_walk(const SplitScratch &sc, int i0, int i1, int k, uint32_t ms, int off_r, int off_a, int off_b, bool with_b, int &qr, int &qa, int &qb, SplitWin *out,
                      int out_cap, int &n, int *s_int) {
#ifdef __CUDA_ARCH__
  // On the device the walk is not walked: every chain anchor learns in parallel where the next cut would be if a cut were made at
  // it; the cuts of the walk are then the anchors reachable from the first cut -- marked by pointer doubling, compacted in index
  // order (the pointers only go forward) -- and every record is written by its own thread.
  const int m = i1 - i0;
  if (m <= 0) return;
  SplitAnchor *ca = reinterpret_cast<SplitAnchor *>(sc.table);             // the chain's anchors in chain order (the table is dead)
  int32_t *mark = reinterpret_cast<int32_t *>(sc.table) + 4 * (size_t)m, *acc = mark + m;
  int32_t *jc = sc.t2_idx, *jn = reinterpret_cast<int32_t *>(sc.t2_km);
  SP_FOR(i, m) { ca[i] = sc.anc[sc.bl[i0 + i]]; mark[i] = 0; }
  if (threadIdx.x == 0) s_int[2] = 0x7fffffff;
  SP_SYNC();
  SP_FOR(i, m) {
    const SplitAnchor me = ca[i];
    if (split_cut_ok(me.r - qr, me.a - qa, me.b - qb, ms)) atomicMin(&s_int[2], i);
    const int pr = me.r + k, pa = me.a + k, pb = me.b + k;
    int j = i + 1;
    for (; j < m; ++j) { const SplitAnchor o = ca[j]; if (split_cut_ok(o.r - pr, o.a - pa, o.b - pb, ms)) break; }
    jc[i] = j < m ? j : -1;
  }
  SP_SYNC();
  const int first = s_int[2];
  SP_SYNC();
  if (first == 0x7fffffff) return;                                          // no anchor ends a record
  if (threadIdx.x == 0) mark[first] = 1;
  SP_SYNC();
  for (int span = 1; span < m; span <<= 1) {
    SP_FOR(i, m) { const int t = jc[i]; if (t >= 0 && mark[i]) mark[t] = 1; jn[i] = t < 0 ? -1 : jc[t]; }
    SP_SYNC();
    int32_t *t = jc; jc = jn; jn = t;
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int count = 0;
    for (int base = 0; base < m; base += 32) {
      const bool on = base + lane < m && mark[base + lane];
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      if (on) acc[count + __popc(bal & ((1u << lane) - 1u))] = base + lane;
      count += __popc(bal);
    }
    if (lane == 0) s_int[3] = count;
  }
  SP_SYNC();
  const int cuts = s_int[3];
  SP_FOR(t, cuts) {
    const SplitAnchor cur = ca[acc[t]];
    int sr = qr, sa = qa, sb = qb;
    if (t > 0) { const SplitAnchor prev = ca[acc[t - 1]]; sr = prev.r + k; sa = prev.a + k; sb = prev.b + k; }
    if (n + t < out_cap) out[n + t] = SplitWin{off_r + sr, cur.r - sr + k, off_a + sa, cur.a - sa + k, with_b ? off_b + sb : 0, with_b ? cur.b - sb + k : 0};
  }
  {
    const SplitAnchor last = ca[acc[cuts - 1]];
    qr = last.r + k; qa = last.a + k; qb = last.b + k;
    n += cuts;
  }
  SP_SYNC(); ST(12);
#else
  (void)s_int;
  for (int i = i0; i < i1; ++i) {
    const SplitAnchor an = sc.anc[sc.bl[i]];
    const int zr = an.r - qr, za = an.a - qa, zb = an.b - qb;
    if (split_cut_ok(zr, za, zb, ms)) {
      if (n < out_cap) out[n] = SplitWin{off_r + qr, zr + k, off_a + qa, za + k, with_b ? off_b + qb : 0, with_b ? zb + k : 0};
      ++n;
      qr = an.r + k; qa = an.a + k; qb = an.b + k;
    }
  }
#endif
}

// The cutting of one job (split(), :175-308) into out[0 ..); returns the number of records, or -1 when out_cap or the anchor arrays
// are too small.  Positions in the records are relative to the three reads given here.  first_call as in the reference: the
// late-start / early-end special cases only at the outer level.
SP_HD int split_job(const SplitScratch &sc, const SplitScratch &sc2, SplitSeq ref, SplitSeq S1, SplitSeq S2, int k, SplitWin *out, int out_cap, int *s_int) {
  const uint32_t min_size = 20;
  const int blen = split_chain(sc, ref, S1, S2, k, min_size, s_int);
  if (blen < 0) return -1;
  if (blen < 1) {   // no anchor: the three reads as one record (:256-261)
    SP_SERIAL { if (out_cap >= 1) out[0] = SplitWin{0, ref.n, 0, S1.n, 0, S2.n}; }
    SP_SYNC();
    return out_cap >= 1 ? 1 : -1;
  }
  // the two special cases run the parallel steps again on sub-strings
  int nout = 0;
  int i0 = 0, pred_r = 0, pred_a = 0, pred_b = 0;
  {
    const int f = sc.bl[0];
    const int sr = sc.anc[f].r + k, sa = sc.anc[f].a + k, sb = sc.anc[f].b + k;   // start_ref / start_S1 / start_S2 lengths (:264-266; substr clamps)
    const int lr = sr < ref.n ? sr : ref.n, la = sa < S1.n ? sa : S1.n, lb = sb < S2.n ? sb : S2.n;
    if ((long long)lb * 2 < lr && (unsigned)(lr - lb) > 200u) {
      // the corrected read starts late: reference and uncorrected prefix are cut on their own (S2 := the reference prefix),
      // with a minimum window of 1.2 x the corrected prefix; the corrected side gets N records and its prefix last (:268-277)
      const SplitSeq pr{ref.s, lr, ref.pk, ref.g}, pa{S1.s, la, S1.pk, S1.g};
      const uint32_t ms = (uint32_t)(1.2 * (double)lb);
      const int bl2 = split_chain(sc2, pr, pa, pr, k, ms, s_int + 4);
      if (bl2 < 0) return -1;
      int n2 = 0, qr = 0, qa = 0, qb = 0;
      if (bl2 >= 1) split_walk(sc2, 0, bl2 - 1, k, ms, 0, 0, 0, false, qr, qa, qb, out, out_cap, n2, s_int);
      SP_SERIAL {
        if (n2 < out_cap) out[n2] = SplitWin{qr, pr.n - qr, qa, pa.n - qa, 0, 0};   // first_call is false there: the rest as it is (:302-306); all of it without an anchor
        // corrected side: n2 times N, then the prefix (N when it is empty) -- generate_dumb_str(n2 + 1, header, start_S2, "")
        for (int i = 0; i <= n2 && i < out_cap; ++i) { out[i].b0 = -1; out[i].bn = 1; }
        if (lb > 0 && n2 < out_cap) { out[n2].b0 = 0; out[n2].bn = lb; }
      }
      SP_SYNC();
      nout = n2 + 1;
      pred_r = sr; pred_a = sa; pred_b = sb;   // (:273-275: the anchor's end, unclamped)
      i0 = 1;
    }
  }
  // the walk (:280-290) and the tail (:292-306)
  split_walk(sc, i0, blen - 1, k, min_size, 0, 0, 0, true, pred_r, pred_a, pred_b, out, out_cap, nout, s_int);
  {
    // substr(pred) of a read that is shorter than pred cannot happen on the chain (positions + k <= length)
    const int er = ref.n - pred_r, ea = S1.n - pred_a, eb = S2.n - pred_b;
    if ((long long)eb * 2 < er && (unsigned)(er - eb) > 200u) {
      // the corrected read ends early (:295-301): the corrected side gets its tail first, then N records
      const SplitSeq pr{ref.s + pred_r, er, ref.pk, ref.g + pred_r}, pa{S1.s + pred_a, ea, S1.pk, S1.g + pred_a};
      const uint32_t ms = (uint32_t)(1.2 * (double)eb);
      const int bl2 = split_chain(sc2, pr, pa, pr, k, ms, s_int + 4);
      if (bl2 < 0) return -1;
      int n2 = nout, qr = 0, qa = 0, qb = 0;
      if (bl2 >= 1) split_walk(sc2, 0, bl2 - 1, k, ms, pred_r, pred_a, 0, false, qr, qa, qb, out, out_cap, n2, s_int);
      SP_SERIAL {
        if (n2 < out_cap) out[n2] = SplitWin{pred_r + qr, er - qr, pred_a + qa, ea - qa, 0, 0};
        for (int i = nout; i <= n2 && i < out_cap; ++i) { out[i].b0 = -1; out[i].bn = 1; }
        if (eb > 0 && nout < out_cap) { out[nout].b0 = pred_b; out[nout].bn = eb; }
      }
      nout = n2 + 1;
    } else {
      SP_SERIAL { if (nout < out_cap) out[nout] = SplitWin{pred_r, er, pred_a, ea, pred_b, eb}; }
      nout = nout + 1;
    }
  }
  SP_SYNC();
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
  const int32_t *order;         // triplets by falling length (the longest jobs start first), or null
  const uint32_t *pk[3];        // the letters of the call at 2 bits each (split_prepack_kernel), or null; letter let[q][i] is letter i - base[q] there
  int64_t base[3];
  // scratch: the job's arrays live in the CTA's dynamic shared memory (smem_words) when they fit, with as many anchors for the
  // sub-calls as for the job itself; a job that does not fit, or whose sub-call finds more anchors than that, runs in the CTA's
  // share of the pool in global memory (cta_words; sub_anchors anchors for the sub-calls)
  uint32_t *pool; uint64_t cta_words; int32_t sub_anchors; uint32_t smem_words;
  // per (triplet, k) job: records, their number (-1: capacity), largest_fragment()
  SplitWin *wins; const int64_t *win_off;   // job (t, ki) at wins[win_off[t] + ki * cap(t)], cap(t) = (win_off[t+1] - win_off[t]) / 4
  int32_t *job_n; uint32_t *job_largest;
  int32_t *counter;
};

// the letters of one kind at 2 bits each, coded like nuc2int (:35-43): the four jobs of a triplet read these words instead of
// coding the bytes four times.  let: 16-byte aligned; one thread per word of 16 letters.
__global__ void __launch_bounds__(256) split_prepack_kernel(const uint8_t *let, int64_t n_letters, uint32_t *pk, int64_t n_words) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    const int64_t t0 = 16 * w;
    if (t0 + 16 <= n_letters) {
      const uint4 q = *reinterpret_cast<const uint4 *>(let + t0);
      const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) { const uint32_t c = (x[j >> 2] >> (8 * (j & 3))) & 0xffu; v |= (c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u) << (2 * j); }
    } else {
      for (int64_t t = t0; t < n_letters; ++t) { const uint32_t c = let[t]; v |= (c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u) << (2 * (int)(t - t0)); }
    }
    pk[w] = v;
  }
}

// persistent CTAs; a job = (triplet, k)
__global__ void __launch_bounds__(1024, 1) split_jobs_kernel(SplitArgs a) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ int s_int[16];
  __shared__ int s_job;
  uint32_t *pool = a.pool + (uint64_t)blockIdx.x * a.cta_words;
  for (;;) {
    if (threadIdx.x == 0) s_job = atomicAdd(a.counter, 1);
    __syncthreads();
    const int64_t job = s_job;
    __syncthreads();
    if (job >= 4 * a.n_triplets) break;
#ifdef SPLIT_TIMING
    if (threadIdx.x == 0) *reinterpret_cast<long long *>(s_int + 12) = clock64();
#endif
    const int64_t t = a.order ? (int64_t)a.order[job >> 2] : job >> 2;
    const int ki = (int)(job & 3), k = 15 - 2 * ki;
    const int64_t slot = 4 * t + ki;
    if (a.status_in[t]) { if (threadIdx.x == 0) { a.job_n[slot] = 0; a.job_largest[slot] = 0; } continue; }
    const SplitSeq ref{a.let[0] + a.off[0][t], (int)(a.off[0][t + 1] - a.off[0][t]), a.pk[0], a.off[0][t] - a.base[0]},
        S1{a.let[1] + a.off[1][t], (int)(a.off[1][t + 1] - a.off[1][t]), a.pk[1], a.off[1][t] - a.base[1]},
        S2{a.let[2] + a.off[2][t], (int)(a.off[2][t + 1] - a.off[2][t]), a.pk[2], a.off[2][t] - a.base[2]};
    const int32_t ma = split_anchor_bound(ref.n, 20);
    const int nb = S2.n > ref.n ? S2.n : ref.n;
    const int64_t cap = (a.win_off[t + 1] - a.win_off[t]) >> 2;
    SplitWin *out = a.wins + a.win_off[t] + (int64_t)ki * cap;
    int n = -2;
    {
      SplitScratch sc, sc2;
      if (split_carve(sc, sc2, s_dyn, a.smem_words, ref.n, S1.n, nb, ma, ma)) n = split_job(sc, sc2, ref, S1, S2, k, out, (int)cap, s_int);
    }
    if (n < 0 && a.cta_words) {   // the same in global memory
      SplitScratch sc, sc2;
      __syncthreads();
      if (split_carve(sc, sc2, pool, a.cta_words, ref.n, S1.n, nb, ma, a.sub_anchors)) n = split_job(sc, sc2, ref, S1, S2, k, out, (int)cap, s_int);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_int[10] = a.header_len[t] + (n > 1 ? 1 : 0);     // largest_fragment(), :158-169 (split_host.hpp)
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine = max(mine, out[i].rn + 1);
    if (mine > 0) atomicMax(&s_int[10], mine);
    __syncthreads();
    if (threadIdx.x == 0) { a.job_n[slot] = n < 0 ? -1 : n; a.job_largest[slot] = (uint32_t)s_int[10]; }
    ST(13);
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

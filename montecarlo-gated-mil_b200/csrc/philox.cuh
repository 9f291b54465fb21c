// philox.cuh — Philox4x32-10 and the dropout-mask convention (mirrors oracle/philox.py).
// Replaces torch's global-generator dropout draws of model.py:281 and :291/:301 by a
// counter-based stream keyed (seed; chunk, patch, sample, bag): masks do not depend on how
// samples or bags are sharded over CTAs or GPUs.
#pragma once
#include <stdint.h>

namespace mcmil {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t ATTN_CHUNK_BASE = 64u;  // ctr.x >= 64: logit-dropout slots

constexpr int PHILOX_MAX_ROUNDS = 10;   // Philox4x32-10 (Random123 / cuRAND / ATen default); 7 is the
                                        // smallest Crush-resistant round count (Salmon et al., SC'11)

struct PhiloxKey { uint32_t k0[PHILOX_MAX_ROUNDS], k1[PHILOX_MAX_ROUNDS]; };

__host__ __device__ inline PhiloxKey philox_key(uint64_t seed) {
  PhiloxKey k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < PHILOX_MAX_ROUNDS; ++r) { k.k0[r] = a; k.k1[r] = b; a += PHILOX_W0; b += PHILOX_W1; }
  return k;
}

// Round keys are pre-expanded on the host (they only depend on the seed) so each round is
// 2 wide multiplies + 2 three-input XORs.
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const PhiloxKey& key) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}
// The same function split for the projection kernel's producers, which draw 16 chunks per thread and
// sample (4 K-slices q x 4 patch rows n): the second-round product M0 * (hi(M1 t) ^ n ^ k0[0]) only
// depends on (sample, row), so it is computed once per sample and row (philox_row_part) and shared by
// the four slices; the q-dependent first / second round terms are common sub-expressions of the four
// rows of a slice.  16 wide multiplies per call instead of 17 (20 without any sharing).
struct PhiloxRowPart { uint32_t p0hi, p0lo; };
__device__ __forceinline__ uint32_t philox_sample_part(uint32_t t) { return (uint32_t)((uint64_t)PHILOX_M1 * t); }  // c1 after round 1
__device__ __forceinline__ PhiloxRowPart philox_row_part(uint32_t n, uint32_t t, const PhiloxKey& key) {
  const uint32_t c0 = (uint32_t)(((uint64_t)PHILOX_M1 * t) >> 32) ^ n ^ key.k0[0];
  const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
  return PhiloxRowPart{(uint32_t)(p0 >> 32), (uint32_t)p0};
}
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_split(uint32_t q, uint32_t bag, uint32_t mt_lo, const PhiloxRowPart& rp,
                                                  const PhiloxKey& key) {
  static_assert(ROUNDS >= 2, "split form needs two rounds");
  const uint64_t m0q = (uint64_t)PHILOX_M0 * q;
  const uint64_t p1 = (uint64_t)PHILOX_M1 * ((uint32_t)(m0q >> 32) ^ bag ^ key.k1[0]);
  uint32_t c0 = (uint32_t)(p1 >> 32) ^ mt_lo ^ key.k0[1];
  uint32_t c1 = (uint32_t)p1;
  uint32_t c2 = rp.p0hi ^ (uint32_t)m0q ^ key.k1[1];
  uint32_t c3 = rp.p0lo;
#pragma unroll
  for (int r = 2; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    const uint64_t q1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(q1 >> 32) ^ c1 ^ key.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
    c1 = (uint32_t)q1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}
// runtime round count (7 or 10) for the non-critical kernels
__device__ __forceinline__ uint4 philox4x32_rt(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKey& key) {
  return rounds == 7 ? philox4x32<7>(c0, c1, c2, c3, key) : philox4x32<10>(c0, c1, c2, c3, key);
}

// keep-bits (bit e = keep element 8q+e) of one feature chunk
__device__ __forceinline__ uint32_t feature_keep8(int rounds, uint32_t q, uint32_t n, uint32_t t, uint32_t bag,
                                                   const PhiloxKey& key, uint32_t thr) {
  const uint4 r = philox4x32_rt(rounds, q, n, t, bag, key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= (((w[i] & 0x7FFFu) >= thr) ? 1u : 0u) << (2 * i);
    bits |= ((((w[i] >> 16) & 0x7FFFu) >= thr) ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

// keep flag of the logit of (bag, t, n), head c
template <int ROUNDS>
__device__ __forceinline__ uint4 attn_words(uint32_t group, uint32_t n, uint32_t t, uint32_t bag,
                                            const PhiloxKey& key) {
  return philox4x32<ROUNDS>(ATTN_CHUNK_BASE + group, n, t, bag, key);
}
__device__ __forceinline__ uint4 attn_words_rt(int rounds, uint32_t group, uint32_t n, uint32_t t, uint32_t bag,
                                               const PhiloxKey& key) {
  return philox4x32_rt(rounds, ATTN_CHUNK_BASE + group, n, t, bag, key);
}
__device__ __forceinline__ bool attn_keep_from(const uint4& r, int c, uint32_t thr) {
  const uint32_t w = (c & 3) == 0 ? r.x : (c & 3) == 1 ? r.y : (c & 3) == 2 ? r.z : r.w;
  return (w & 0x7FFFu) >= thr;
}

}  // namespace mcmil

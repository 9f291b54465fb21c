// philox.cuh — Philox4x32-10 and the dropout-mask convention (mirrors oracle/philox.py).
// Replaces torch's global-generator dropout draws of model.py:281 and :291/:301 by a
// counter-based stream keyed on (seed; feature chunk, patch, sample, bag): masks do not depend on
// how samples or bags are sharded over CTAs or GPUs.
//
// Feature (t, n, l) draws a 15-bit uniform lane = ((P << 8) | R) & 0x7fff and is kept iff lane >= thr:
//   P = byte 8*((n>>2)&1) + l%8        of philox(l/8, n & ~4, t, bag)                 "primary" call
//   R = byte 4*((n>>2)&3) + (l/64 >> 1) of philox(128 + (l/64 & 1)*8 + (l%64)/8, n & ~12, t, bag)   "refinement" call
// One primary call serves 16 elements (8 features of the two rows n, n^4); the refinement byte only
// decides when the 7 primary bits equal the top bits of the threshold (1 draw in 128) and is shared
// by the 8 features of a (row, chunk): 9 Philox calls per 128 features instead of 16 with one 16-bit
// lane per feature, same 15-bit resolution of p; the shared byte couples two features of a chunk
// only when both hit that 1-in-128 case (pairwise correlation ~1e-4).
#pragma once
#include <stdint.h>

namespace mcmil {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t ATTN_CHUNK_BASE = 64u;  // ctr.x in [64, 128): logit-dropout slots
constexpr uint32_t REF_CHUNK_BASE = 128u;  // ctr.x in [128, 144): refinement bytes of the feature masks

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
// runtime round count (7 or 10) for the non-critical kernels
__device__ __forceinline__ uint4 philox4x32_rt(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKey& key) {
  return rounds == 7 ? philox4x32<7>(c0, c1, c2, c3, key) : philox4x32<10>(c0, c1, c2, c3, key);
}

// keep-bits (bit e = keep element 8q+e) of one feature chunk: the convention spelled out element by
// element (export / fp32 cross-check kernels; the tcgen05 producers evaluate it 16 elements per call)
__device__ __forceinline__ uint32_t feature_keep8(int rounds, uint32_t q, uint32_t n, uint32_t t, uint32_t bag,
                                                   const PhiloxKey& key, uint32_t thr) {
  const uint4 pr = philox4x32_rt(rounds, q, n & ~4u, t, bag, key);
  const uint4 rf = philox4x32_rt(rounds, REF_CHUNK_BASE + ((q >> 3) & 1u) * 8u + (q & 7u), n & ~12u, t, bag, key);
  const uint32_t wa = (n & 4u) ? pr.z : pr.x, wb = (n & 4u) ? pr.w : pr.y;
  const uint32_t ri = (n >> 2) & 3u;
  const uint32_t rw = ri == 0 ? rf.x : ri == 1 ? rf.y : ri == 2 ? rf.z : rf.w;
  const uint32_t rb = (rw >> (8u * (q >> 4))) & 0xFFu;
  uint32_t bits = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t pb = ((e < 4 ? wa : wb) >> (8 * (e & 3))) & 0xFFu;
    bits |= ((((pb << 8) | rb) & 0x7FFFu) >= thr ? 1u : 0u) << e;
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

// reduce.cu — the reductions behind the projection: per-head softmax over the variable-length
// patch axis (model.py:305), the attention-weighted classifier contraction (model.py:308-316,
// as sum_n A[t,c,n]*score[t,c,n]), and the Welford mean / M2 over the T MC samples of the class
// probabilities (infer.py:195, net_utils.py:207-208) and of the attention (infer.py:212-219).
// Warp-shuffle kernels, 16-byte accesses along the patch axis, no float atomics (run-to-run deterministic).
//
// Input: the logit / score planes [T][C][Rp] of the projection kernel; every bag starts at a multiple of 32
// columns (Plan::d_pcol), so a bag's row segment is 128-byte aligned.  Two launches: rows (`softmax_rows_*`: max,
// sum exp and the pooled classifier logit of every (bag, t, c) row), then columns (`welford_cols_kernel`: attention
// values and their Welford statistics over the samples, MC samples split over CTAs when there are few tiles).
// (A one-launch variant — a cluster of 8 CTAs per bag with the slabs in shared memory and the row partials exchanged
// through distributed shared memory — was built and measured in round 2: 19 us per single bag against 10.7 us for
// the two launches, because it runs on 8 SMs instead of 148; removed, profiles/r2_experiments.md.)
#include "internal.h"

namespace mcmil {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exp / reciprocal without the range fix-ups of __expf / __fdividef (arguments here are <= 0 resp. small positive
// integers; a denormal result flushing to zero is irrelevant): 2 instructions instead of 6
__device__ __forceinline__ float fast_exp(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f)); return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ void grid_dep_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");               // results of the previous kernel in the stream
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Chan merge of two (count, mean, M2) partials
__device__ __forceinline__ void wf_merge(float& n_a, float& mu, float& q, float n_b, float mb, float qb) {
  if (n_b <= 0.f) return;
  const float n_ab = n_a + n_b, dlt = mb - mu;
  mu += dlt * __fdividef(n_b, n_ab);
  q += qb + dlt * dlt * __fdividef(n_a * n_b, n_ab);
  n_a = n_ab;
}
// softmax_c of one sample's logits (C <= MAXC, register-only: no runtime-indexed arrays)
__device__ __forceinline__ void softmax_classes(const float* y, int C, float (&p)[MAXC]) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { p[c] = c < C ? y[c] : -INFINITY; mx = fmaxf(mx, p[c]); }
  float z = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { p[c] = fast_exp(p[c] - mx); z += p[c]; }      // exp(-inf) = 0 for c >= C
  const float iz = 1.0f / z;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) p[c] *= iz;
}
// mean / M2 over t of softmax_c(y[t][:]) by one warp (two passes: the values are tiny); y: T x C, any address space
__device__ __forceinline__ void prob_stats_warp(const float* y, int T, int C, int lane, float* prob_mean, float* prob_m2) {
  float s[MAXC] = {0.f, 0.f, 0.f, 0.f};
  for (int t = lane; t < T; t += 32) {
    float p[MAXC];
    softmax_classes(y + t * C, C, p);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) s[c] += p[c];
  }
  float mean[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) mean[c] = warp_sum(s[c]) / (float)T;
  float q[MAXC] = {0.f, 0.f, 0.f, 0.f};
  for (int t = lane; t < T; t += 32) {
    float p[MAXC];
    softmax_classes(y + t * C, C, p);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { const float dlt = p[c] - mean[c]; q[c] = fmaf(dlt, dlt, q[c]); }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const float qq = warp_sum(q[c]);
    if (lane == 0 && c < C) { prob_mean[c] = mean[c]; if (prob_m2) prob_m2[c] = qq; }
  }
}

// ================================================================================== generic path, launch 1: rows
constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;

// WPR warps per (bag, t, c) row, ROW_WARPS / WPR rows per CTA, any row length.  A warp walks its share of the row in
// chunks of 128 * V4 patches (V4 float4 per lane and plane, all 2 * V4 loads of a chunk issued before the first use);
// max and exp / sum of a chunk run on the registers and chunks are merged with the warp's running (max, sum, pooled
// score) on the fly, so each plane is read from DRAM exactly once and no warp ever waits for another one before the
// end of the row.
//   WPR = 1 (many rows: >= 2048): a warp owns a row.  LOOP = false: every row fits one chunk (no loop-carried
//     state: 80 registers, 3 CTAs per SM); LOOP = true: 98 registers, 2 CTAs per SM.
//   WPR = 8 (few rows: one bag, or long rows): the CTA's warps take the row's chunks round-robin and their running
//     statistics are merged through shared memory once, after one __syncthreads.  V4 is chosen by the launcher so
//     that a row of up to 8 * 128 * V4 patches is one chunk per warp.  (Round 1/2 kept a whole 16384-patch chunk in
//     the registers of one CTA: 157 registers, ONE CTA per SM, two block-wide reductions per row — 75 % of the copy
//     peak at config 4, and 64 exp of -inf per thread on the 1024-patch rows of a single-bag call.)
// Block 0 also clears the arrival counters of the column kernel.
constexpr int ROWW_V4 = 8;
template <bool LOOP, int WPR, int V4>
__global__ void __launch_bounds__(ROW_THREADS, (LOOP && V4 == 8) ? 2 : 3)
softmax_rows_warp_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                         const int32_t* __restrict__ cu, const int32_t* __restrict__ pcol, int n_bags, int T, int C,
                         int Rp, float2* __restrict__ rowstat, float* __restrict__ Y, int* __restrict__ wcount,
                         int n_wcount) {
  static_assert(WPR == 1 || WPR == ROW_WARPS, "a warp or the whole CTA per row");
  static_assert(LOOP || WPR == 1, "the one-chunk form is for warp-per-row launches");
  __shared__ float red[3][ROW_WARPS];
  grid_dep_sync();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_wcount; i += ROW_THREADS) wcount[i] = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = WPR == 1 ? 0 : warp;                        // this warp's position among the row's warps
  const long long row = WPR == 1 ? (long long)blockIdx.x * ROW_WARPS + warp : (long long)blockIdx.x;
  if (row >= (long long)n_bags * T * C) return;               // (WPR = 8: uniform over the CTA)
  const int c = (int)(row % C);
  const int t = (int)((row / C) % T);
  const int b = (int)(row / ((long long)C * T));
  const int n = cu[b + 1] - cu[b];
  const size_t off = ((size_t)t * C + c) * Rp + pcol[b];
  const float4* lg = reinterpret_cast<const float4*>(logits + off);
  const float4* sc = reinterpret_cast<const float4*>(scores + off);
  float M = -INFINITY, Z = 0.f, Yv = 0.f;                     // running statistics of this warp's chunks (warp-uniform)
  for (int base = sub * 128 * V4; base < (LOOP ? n : 1); base += WPR * 128 * V4) {
    float4 v[V4], w[V4];
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const int i4 = base / 4 + lane + 32 * k;
      if (4 * i4 < n) { v[k] = __ldg(lg + i4); w[k] = __ldg(sc + i4); }
      else { v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); w[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const int col = base + 4 * (lane + 32 * k);             // the bag's last float4 may reach into the plane padding
      if (col + 1 >= n) { v[k].y = -INFINITY; w[k].y = 0.f; }
      if (col + 2 >= n) { v[k].z = -INFINITY; w[k].z = 0.f; }
      if (col + 3 >= n) { v[k].w = -INFINITY; w[k].w = 0.f; }
      m = fmaxf(fmaxf(m, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
    m = warp_max(m);
    float z = 0.f, y = 0.f;
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const float e0 = fast_exp(v[k].x - m), e1 = fast_exp(v[k].y - m);      // exp(-inf) = 0 for the padding
      const float e2 = fast_exp(v[k].z - m), e3 = fast_exp(v[k].w - m);
      z += (e0 + e1) + (e2 + e3);
      y = fmaf(e0, w[k].x, y); y = fmaf(e1, w[k].y, y); y = fmaf(e2, w[k].z, y); y = fmaf(e3, w[k].w, y);
    }
    z = warp_sum(z); y = warp_sum(y);
    const float Mn = fmaxf(M, m);
    const float fa = fast_exp(M - Mn), fb = fast_exp(m - Mn);               // exp(-inf) = 0 on the first chunk
    Z = fmaf(Z, fa, z * fb);
    Yv = fmaf(Yv, fa, y * fb);
    M = Mn;
  }
  if constexpr (WPR > 1) {
    // merge the warps' statistics in warp order (a warp without a chunk holds (-inf, 0, 0) and drops out)
    if (lane == 0) { red[0][warp] = M; red[1][warp] = Z; red[2][warp] = Yv; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    M = red[0][0]; Z = red[1][0]; Yv = red[2][0];
#pragma unroll
    for (int q = 1; q < ROW_WARPS; ++q) {
      const float m = red[0][q];
      if (m == -INFINITY) continue;
      const float Mn = fmaxf(M, m);
      const float fa = fast_exp(M - Mn), fb = fast_exp(m - Mn);
      Z = fmaf(Z, fa, red[1][q] * fb);
      Yv = fmaf(Yv, fa, red[2][q] * fb);
      M = Mn;
    }
  }
  if (lane == 0) {
    const float inv = 1.0f / Z;
    rowstat[((size_t)c * n_bags + b) * T + t] = make_float2(M, inv);   // [C][n_bags][T]: a row's samples are contiguous
    Y[((size_t)b * T + t) * C + c] = Yv * inv;
  }
}

// ================================================================================== generic path, launch 2: columns
constexpr int COL_LANES = 32;
constexpr int COL_VEC = 4;        // patches per lane: one 16-byte load per sample
constexpr int COL_WARPS = 8;
constexpr int COL_THREADS = COL_LANES * COL_WARPS;
#ifndef MCMIL_COL_TPC
#define MCMIL_COL_TPC 2           // projection tiles (128 patches each) per column CTA
#endif
#ifndef MCMIL_COL_UNROLL
#define MCMIL_COL_UNROLL 4        // samples a warp has in flight (independent 512-byte loads)
#endif
constexpr int COL_TPC = MCMIL_COL_TPC;
constexpr int COL_TGROUPS = COL_WARPS / COL_TPC;   // the CTA's samples are strided over this many warps per tile
constexpr int COL_UNROLL = MCMIL_COL_UNROLL;
constexpr int COL_MAX_SPLIT = 16;
constexpr int COL_KPT = (COL_TPC * TILE_ROWS + COL_THREADS - 1) / COL_THREADS;   // patches per thread in the final merge
static_assert(COL_LANES * COL_VEC == TILE_ROWS && (COL_TPC == 1 || COL_TPC == 2 || COL_TPC == 4 || COL_TPC == 8), "column CTA shape");
int col_tiles_per_cta() { return COL_TPC; }

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// grid (n_cblk + bag_blocks, C, wsplit).
// blockIdx.x < n_cblk: CTA = (block of up to COL_TPC consecutive 128-patch tiles of one bag, head c, sample group z):
//   samples [T z / wsplit, T (z+1) / wsplit).  Warp w works on tile w % COL_TPC of the block and takes the group's
//   samples w / COL_TPC, + COL_TGROUPS, ...: every lane reads 4 adjacent patches of a sample with one 16-byte load
//   (the warps of one sample row read COL_TPC * 512 contiguous bytes), COL_UNROLL samples in flight per warp together
//   with their (max, 1 / sum), no shared-memory staging; A = exp(l - max) / sum and the Welford update run as packed
//   fp32x2 operations (two patches per instruction).  The partial (count, mean, M2) of the sample groups are merged
//   with Chan's formula in group order.
//   wsplit > 1 (few tiles, many samples: one large bag): every CTA writes its partial to the workspace and the
//   LAST CTA of the (block, head) to arrive (one integer atomic per CTA) merges the groups in the order 0..wsplit-1,
//   so the result does not depend on the arrival order.  Optionally stores A.
// blockIdx.x >= n_cblk (y = z = 0): one warp per bag: mean / M2 over t of softmax_c(Y[bag][t][:])
#ifndef MCMIL_COL_MINB
#define MCMIL_COL_MINB 4     // CTAs per SM the column kernel is compiled for (64 registers)
#endif
template <bool HAS_A>
__global__ void __launch_bounds__(COL_THREADS, MCMIL_COL_MINB)
welford_cols_kernel(const float* __restrict__ logits, const float2* __restrict__ rowstat,
                    const int4* __restrict__ cblk, const float* __restrict__ Y,
                    int n_bags, int T, int C, int R, int Rp, int n_cblk, int wsplit,
                    float2* __restrict__ wpart, int* __restrict__ wcount,
                    float* __restrict__ A, float* __restrict__ attn_mean, float* __restrict__ attn_m2,
                    float* __restrict__ prob_mean, float* __restrict__ prob_m2) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  grid_dep_sync();                                           // rowstat / Y of the row kernel
  if ((int)blockIdx.x >= n_cblk) {
    if (blockIdx.y != 0 || blockIdx.z != 0 || prob_mean == nullptr) return;
    const int b = ((int)blockIdx.x - n_cblk) * COL_WARPS + warp;
    if (b >= n_bags) return;
    prob_stats_warp(Y + (size_t)b * T * C, T, C, lane, prob_mean + b * C, prob_m2 ? prob_m2 + b * C : nullptr);
    return;
  }
  __shared__ float s_mean[COL_TGROUPS][COL_TPC * TILE_ROWS], s_m2[COL_TGROUPS][COL_TPC * TILE_ROWS];
  __shared__ int s_cnt[COL_WARPS];
  __shared__ int s_last;
  const int4 blk = cblk[blockIdx.x];                         // one load: (plane column, packed row, bag, patches) of the block
  const int cg = warp % COL_TPC, grp = warp / COL_TPC;
  const int pcol_b = blk.x, row_b = blk.y, bag = blk.z, ncols_b = blk.w;
  const int c = blockIdx.y, z = blockIdx.z;
  const int t_lo = (int)((long long)T * z / wsplit), t_hi = (int)((long long)T * (z + 1) / wsplit);
  // (the last 16-byte load of the block may reach into the bag's plane padding: loaded, its results never stored)
  const int col_w = cg * TILE_ROWS + lane * COL_VEC;         // this lane's first patch within the block
  const float* plane = logits + (size_t)c * Rp + pcol_b + col_w;
  const size_t tstride = (size_t)C * Rp;
  const float2* rs_row = rowstat + ((size_t)c * n_bags + bag) * T;
  uint64_t mean01 = 0ull, mean23 = 0ull, q01 = 0ull, q23 = 0ull;
  float cntf = 0.f;
  constexpr float LOG2E = 1.4426950408889634f;
  // one sample of this lane's 4 patches: a = 2^(l log2e - max log2e) / sum, then Welford, two patches per instruction
  auto push = [&](const float4& l4, const float2& rs, int tu) {
    const float nm = -rs.x * LOG2E;
    const uint64_t k2 = f2_pack(LOG2E, LOG2E), nm2 = f2_pack(nm, nm), inv2 = f2_pack(rs.y, rs.y);
    float x0, x1, x2, x3;
    f2_unpack(f2_fma(f2_pack(l4.x, l4.y), k2, nm2), x0, x1);
    f2_unpack(f2_fma(f2_pack(l4.z, l4.w), k2, nm2), x2, x3);
    float e0, e1, e2, e3;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(x2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(x3));
    const uint64_t a01 = f2_mul(f2_pack(e0, e1), inv2), a23 = f2_mul(f2_pack(e2, e3), inv2);
    if constexpr (HAS_A) {
      float a0, a1, a2, a3;
      f2_unpack(a01, a0, a1);
      f2_unpack(a23, a2, a3);
      float* dst = A + ((size_t)tu * C + c) * R + row_b + col_w;
      const int left = ncols_b - col_w;
      dst[0] = a0;
      if (left > 1) dst[1] = a1;
      if (left > 2) dst[2] = a2;
      if (left > 3) dst[3] = a3;
    }
    cntf += 1.0f;
    const float ic = fast_rcp(cntf);
    const uint64_t ic2 = f2_pack(ic, ic);
    const uint64_t d01 = f2_sub(a01, mean01), d23 = f2_sub(a23, mean23);
    mean01 = f2_fma(d01, ic2, mean01);
    mean23 = f2_fma(d23, ic2, mean23);
    q01 = f2_fma(d01, f2_sub(a01, mean01), q01);
    q23 = f2_fma(d23, f2_sub(a23, mean23), q23);
  };
  const int n_my = col_w < ncols_b ? (t_hi - t_lo - grp + COL_TGROUPS - 1) / COL_TGROUPS : 0;   // this warp's samples
  if (n_my > 0) {
    // pointer-increment addressing: the COL_UNROLL loads of a batch only differ by compile-time multiples of one
    // 64-bit step, so they are issued back to back (index arithmetic between the loads serialised them on the
    // scoreboard: profiles/r2_experiments.md)
    const float4* lp = reinterpret_cast<const float4*>(plane + (size_t)(t_lo + grp) * tstride);
    const size_t step = (size_t)COL_TGROUPS * (tstride / 4);           // float4 units (Rp is a multiple of 32)
    const float2* rp = rs_row + t_lo + grp;
    int tu = t_lo + grp, i = 0;
    for (; i + COL_UNROLL <= n_my; i += COL_UNROLL) {
      float4 l4[COL_UNROLL];
      float2 rs[COL_UNROLL];
#pragma unroll
      for (int u = 0; u < COL_UNROLL; ++u) { l4[u] = __ldg(lp + u * step); rs[u] = __ldg(rp + u * COL_TGROUPS); }
      lp += COL_UNROLL * step;
      rp += COL_UNROLL * COL_TGROUPS;
#pragma unroll
      for (int u = 0; u < COL_UNROLL; ++u) push(l4[u], rs[u], tu + u * COL_TGROUPS);
      tu += COL_UNROLL * COL_TGROUPS;
    }
    for (; i < n_my; ++i) {
      const float4 l4 = __ldg(lp);
      const float2 rs = __ldg(rp);
      lp += step; rp += COL_TGROUPS;
      push(l4, rs, tu);
      tu += COL_TGROUPS;
    }
  }
  const int cnt = n_my;
  {
    float m0, m1, m2_, m3, v0, v1, v2, v3;
    f2_unpack(mean01, m0, m1); f2_unpack(mean23, m2_, m3);
    f2_unpack(q01, v0, v1); f2_unpack(q23, v2, v3);
    *reinterpret_cast<float4*>(&s_mean[grp][col_w]) = make_float4(m0, m1, m2_, m3);
    *reinterpret_cast<float4*>(&s_m2[grp][col_w]) = make_float4(v0, v1, v2, v3);
    if (lane == 0) s_cnt[warp] = cnt;                         // (0 for a warp whose tile does not exist in this block)
  }
  __syncthreads();
  float mu[COL_KPT], q[COL_KPT];
#pragma unroll
  for (int k = 0; k < COL_KPT; ++k) {
    const int col = threadIdx.x + k * COL_THREADS;
    mu[k] = 0.f; q[k] = 0.f;
    float n_a = 0.f;
    if (col < ncols_b) {
#pragma unroll
      for (int g = 0; g < COL_TGROUPS; ++g)
        wf_merge(n_a, mu[k], q[k], (float)s_cnt[g * COL_TPC + col / TILE_ROWS], s_mean[g][col], s_m2[g][col]);
    }
  }
  if (wsplit == 1) {
#pragma unroll
    for (int k = 0; k < COL_KPT; ++k) {
      const int col = threadIdx.x + k * COL_THREADS;
      if (col < ncols_b) {
        if (attn_mean) attn_mean[(size_t)c * R + row_b + col] = mu[k];
        if (attn_m2) attn_m2[(size_t)c * R + row_b + col] = q[k];
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < COL_KPT; ++k) {
    const int col = threadIdx.x + k * COL_THREADS;
    if (col < ncols_b) wpart[((size_t)z * C + c) * Rp + pcol_b + col] = make_float2(mu[k], q[k]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&wcount[blockIdx.x * C + c], 1) == wsplit - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < COL_KPT; ++k) {
    const int col = threadIdx.x + k * COL_THREADS;
    if (col < ncols_b) {
      float m = 0.f, v = 0.f, n_a = 0.f;
      for (int g = 0; g < wsplit; ++g) {
        const float2 pk = __ldcg(&wpart[((size_t)g * C + c) * Rp + pcol_b + col]);
        const int ng = (int)((long long)T * (g + 1) / wsplit) - (int)((long long)T * g / wsplit);
        wf_merge(n_a, m, v, (float)ng, pk.x, pk.y);
      }
      if (attn_mean) attn_mean[(size_t)c * R + row_b + col] = m;
      if (attn_m2) attn_m2[(size_t)c * R + row_b + col] = v;
    }
  }
}

int welford_split(int n_cblk, int C, int T) {
  // Sample groups per (column block, head).  The kernel runs up to 4 CTAs per SM (592 CTA slots on a B200): with at
  // least one full wave of blocks there is nothing to gain (config 2: splitting the 100 samples in two made the kernel
  // 30 % slower — partial writes, atomics, a second tail); with fewer blocks (one bag, or one large bag) the samples
  // are split so that the grid gives every SM about three CTAs, at least 16 samples per group.  (Config 4, 128
  // blocks: 2 / 3 / 4 / 8 / 9 / 16 groups measured 35.4 / 34.2 / 37.5-40.0 / 39.2 / 38.7 / 42.7 us — a CTA's prologue,
  // partial write and arrival atomic are not free, and a plain read of the same 131 MB needs ~27-30 us:
  // profiles/r2_experiments.md.)
  const long long ctas = (long long)n_cblk * C, slots = 3 * 148;
  if (ctas >= 4 * 148) return 1;
#ifdef MCMIL_EXP_WSPLIT       // timing experiments: a fixed split below one wave
  return (MCMIL_EXP_WSPLIT) <= T / 16 ? (MCMIL_EXP_WSPLIT) : (T / 16 < 1 ? 1 : T / 16);
#endif
  long long s = slots / ctas;
  if (s > COL_MAX_SPLIT) s = COL_MAX_SPLIT;
  if (s > T / 16) s = T / 16;
  return s < 1 ? 1 : (int)s;
}

cudaError_t launch_reduce(const Plan& p, const float* logits, const float* scores, uint8_t* workspace,
                          float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean,
                          float* attn_m2, cudaStream_t st, int* launches) {
  const int32_t* cu = p.d_cu;
  const int32_t* pcol = p.d_pcol;
  int n_bags = p.n_bags, T = p.T, C = p.C, R = p.R, Rp = p.Rp;
  float2* rowstat = reinterpret_cast<float2*>(workspace + p.off_rowstat);
  float2* wpart = reinterpret_cast<float2*>(workspace + p.off_wpart);
  int* wcount = reinterpret_cast<int*>(workspace + p.off_wcount);
  int n_wcount = p.wsplit > 1 ? p.n_cblk * p.C : 0;
  {
    const long long rows = (long long)p.n_bags * p.T * p.C;
    const unsigned warp_grid = (unsigned)((rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32));
    cudaError_t e;
    // one warp per row when there are enough rows to fill the GPU that way and the rows are not so long that a
    // CTA per row streams them better; otherwise a CTA per row with the smallest chunk that covers the longest bag
    auto launch = [&](auto kernel, unsigned grid) {
      PdlLaunch L(dim3(grid), dim3(ROW_THREADS), 0, st, p.sm_limit == 0);
      return cudaLaunchKernelEx(&L.cfg, kernel, logits, scores, cu, pcol, n_bags, T, C, Rp, rowstat, Y, wcount, n_wcount);
    };
    if (p.max_n <= 128 * ROWW_V4 && rows >= 2048) e = launch(softmax_rows_warp_kernel<false, 1, ROWW_V4>, warp_grid);
    else if (p.max_n <= 8192 && rows >= 2048) e = launch(softmax_rows_warp_kernel<true, 1, ROWW_V4>, warp_grid);
    else if (p.max_n <= ROW_WARPS * 128 * 1) e = launch(softmax_rows_warp_kernel<true, ROW_WARPS, 1>, (unsigned)rows);
    else if (p.max_n <= ROW_WARPS * 128 * 2) e = launch(softmax_rows_warp_kernel<true, ROW_WARPS, 2>, (unsigned)rows);
    else if (p.max_n <= ROW_WARPS * 128 * 4) e = launch(softmax_rows_warp_kernel<true, ROW_WARPS, 4>, (unsigned)rows);
    else e = launch(softmax_rows_warp_kernel<true, ROW_WARPS, ROWW_V4>, (unsigned)rows);
    if (e != cudaSuccess) return e;
  }
  if (launches) ++*launches;
  const int bag_blocks = (p.n_bags + COL_WARPS - 1) / COL_WARPS;
  {
    PdlLaunch L(dim3(p.n_cblk + bag_blocks, p.C, p.wsplit), dim3(COL_THREADS), 0, st, p.sm_limit == 0);
    const float2* rs = rowstat;
    const int4* cblk = p.d_cblk;
    const float* Yc = Y;
    int n_cblk = p.n_cblk, wsplit = p.wsplit;
    cudaError_t e = A != nullptr
        ? cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<true>, logits, rs, cblk, Yc, n_bags, T, C, R, Rp, n_cblk, wsplit,
                             wpart, wcount, A, attn_mean, attn_m2, prob_mean, prob_m2)
        : cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<false>, logits, rs, cblk, Yc, n_bags, T, C, R, Rp, n_cblk, wsplit,
                             wpart, wcount, A, attn_mean, attn_m2, prob_mean, prob_m2);
    if (e != cudaSuccess) return e;
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ Welford sum-form (multi-GPU merge)
__global__ void welford_pack_kernel(const float* __restrict__ mean, const float* __restrict__ m2, double count,
                                    int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) out[0] = count;
  if (i < n) {
    const double mu = (double)mean[i];
    out[1 + i] = count * mu;
    out[1 + n + i] = (double)m2[i] + count * mu * mu;
  }
}
__global__ void welford_unpack_kernel(const double* __restrict__ in, int n, float* __restrict__ mean,
                                      float* __restrict__ m2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double cnt = in[0];
    const double mu = in[1 + i] / cnt;
    mean[i] = (float)mu;
    const double v = in[1 + n + i] - cnt * mu * mu;
    m2[i] = (float)(v > 0.0 ? v : 0.0);
  }
}
cudaError_t launch_welford_pack(const float* mean, const float* m2, double count, int n, double* packed,
                                cudaStream_t st) {
  welford_pack_kernel<<<(n + 255) / 256 + 1, 256, 0, st>>>(mean, m2, count, n, packed);
  return cudaGetLastError();
}
cudaError_t launch_welford_unpack(const double* packed, int n, float* mean, float* m2, cudaStream_t st) {
  welford_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(packed, n, mean, m2);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ auxiliary loss (model.py:318-326, 405-426)
// One CTA per (bag, sample): d = || A[t,pos,:] - A[t,neg,:] + eps ||_2 over the bag's patches
// (F.pairwise_distance adds eps to the difference), loss = scale * (positive ? max(margin - d, 0) : d).
// Fixed-order tree reduction: bit-deterministic.
__global__ void __launch_bounds__(256) aux_pairwise_kernel(const float* __restrict__ A, const int* __restrict__ cu, int T,
                                                           int C, int R, int pos, int neg, int is_positive, float margin,
                                                           float scale, float eps, float* __restrict__ loss) {
  const int b = blockIdx.x, t = blockIdx.y;
  const int r0 = cu[b], r1 = cu[b + 1];
  const float* ap = A + ((size_t)t * C + pos) * R;
  const float* an = A + ((size_t)t * C + neg) * R;
  float acc = 0.f;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const float d = ap[r] - an[r] + eps;
    acc = fmaf(d, d, acc);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    const float d = sqrtf(s);
    loss[(size_t)b * T + t] = scale * (is_positive ? fmaxf(margin - d, 0.f) : d);
  }
}
cudaError_t launch_aux_pairwise(const Plan& p, const float* A, int pos, int neg, int is_positive, float margin,
                                float scale, float eps, float* loss, cudaStream_t st) {
  aux_pairwise_kernel<<<dim3(p.n_bags, p.T), 256, 0, st>>>(A, p.d_cu, p.T, p.C, p.R, pos, neg, is_positive, margin, scale,
                                                          eps, loss);
  return cudaGetLastError();
}

}  // namespace mcmil

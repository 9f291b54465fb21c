// reduce.cu — the reductions behind the projection: per-head softmax over the variable-length
// patch axis (model.py:305), the attention-weighted classifier contraction (model.py:308-316,
// as sum_n A[t,c,n]*score[t,c,n]), and the Welford mean / M2 over the T MC samples of the class
// probabilities (infer.py:195, net_utils.py:207-208) and of the attention (infer.py:212-219).
// Warp-shuffle kernels, 16-byte accesses along the patch axis, no float atomics (run-to-run deterministic).
//
// Input: the logit / score planes [T][C][Rp] of the projection kernel; every bag starts at a multiple of 32
// columns (Plan::d_pcol), so a bag's row segment is 128-byte aligned.  Two paths:
//   * one launch (`fused_bag_reduce_kernel`): a cluster of 8 CTAs per bag stages the bag's slabs of both planes
//     in shared memory, exchanges the per-row softmax partials through distributed shared memory and finishes
//     every output of the bag — the latency path of the reference's bs == 1 serving loop (infer.py:187-196);
//   * two launches (`softmax_rows_*` then `welford_cols_kernel`) for bags whose slabs do not fit 227 KB or for
//     large batches: rows by warps / CTAs, then columns with the MC samples split over CTAs.
#include <mutex>
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
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void grid_dep_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");               // results of the previous kernel in the stream
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Welford update / Chan merge, shared by both paths
__device__ __forceinline__ void wf_push(float& mean, float& m2, float a, float inv_cnt) {
  const float d = a - mean;
  mean = fmaf(d, inv_cnt, mean);
  m2 = fmaf(d, a - mean, m2);
}
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

// One WARP per (bag, t, c) row, eight rows per CTA, any row length: the row is walked in chunks of 1024 patches
// (8 float4 per lane and plane, all 16 loads of a chunk issued before the first use); max and exp / sum of a chunk run
// on the registers, chunks are merged with the running (max, sum) on the fly, so each plane is read from DRAM exactly
// once.  Two CTAs per SM (98 registers) keep the loads of other warps in flight while one warp reduces.  Block 0 also
// clears the arrival counters of the column kernel.
constexpr int ROWW_V4 = 8;
__global__ void __launch_bounds__(ROW_THREADS, 2)
softmax_rows_warp_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                         const int32_t* __restrict__ cu, const int32_t* __restrict__ pcol, int n_bags, int T, int C,
                         int Rp, float2* __restrict__ rowstat, float* __restrict__ Y, int* __restrict__ wcount,
                         int n_wcount) {
  grid_dep_sync();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_wcount; i += ROW_THREADS) wcount[i] = 0;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= (long long)n_bags * T * C) return;
  const int c = (int)(row % C);
  const int t = (int)((row / C) % T);
  const int b = (int)(row / ((long long)C * T));
  const int n = cu[b + 1] - cu[b];
  const size_t off = ((size_t)t * C + c) * Rp + pcol[b];
  const float4* lg = reinterpret_cast<const float4*>(logits + off);
  const float4* sc = reinterpret_cast<const float4*>(scores + off);
  float M = -INFINITY, Z = 0.f, Yv = 0.f;                     // running row statistics (warp-uniform)
  for (int base = 0; base < n; base += 128 * ROWW_V4) {
    float4 v[ROWW_V4], w[ROWW_V4];
#pragma unroll
    for (int k = 0; k < ROWW_V4; ++k) {
      const int i4 = base / 4 + lane + 32 * k;
      if (4 * i4 < n) { v[k] = __ldg(lg + i4); w[k] = __ldg(sc + i4); }
      else { v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); w[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < ROWW_V4; ++k) {
      const int col = base + 4 * (lane + 32 * k);             // the bag's last float4 may reach into the plane padding
      if (col + 1 >= n) { v[k].y = -INFINITY; w[k].y = 0.f; }
      if (col + 2 >= n) { v[k].z = -INFINITY; w[k].z = 0.f; }
      if (col + 3 >= n) { v[k].w = -INFINITY; w[k].w = 0.f; }
      m = fmaxf(fmaxf(m, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
    m = warp_max(m);
    float z = 0.f, y = 0.f;
#pragma unroll
    for (int k = 0; k < ROWW_V4; ++k) {
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
  if (lane == 0) {
    const float inv = 1.0f / Z;
    rowstat[((size_t)c * n_bags + b) * T + t] = make_float2(M, inv);   // [C][n_bags][T]: a row's samples are contiguous
    Y[((size_t)b * T + t) * C + c] = Yv * inv;
  }
}

// Long rows: one CTA per (bag, t, c) row.  A chunk of up to 16384 patches of both planes lives in registers
// (16 float4 per thread and plane, all loads issued before the first use), so a row of that length is read from
// DRAM exactly once; longer rows take several chunks, merged with the running (max, sum) on the fly.
constexpr int ROWC_V4 = 16;
__global__ void __launch_bounds__(ROW_THREADS)
softmax_rows_cta_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                        const int32_t* __restrict__ cu, const int32_t* __restrict__ pcol, int n_bags, int T, int C,
                        int Rp, float2* __restrict__ rowstat, float* __restrict__ Y, int* __restrict__ wcount,
                        int n_wcount) {
  __shared__ float red[3][ROW_THREADS / 32];
  grid_dep_sync();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_wcount; i += ROW_THREADS) wcount[i] = 0;
  const int c = blockIdx.x % C;
  const int t = (blockIdx.x / C) % T;
  const int b = blockIdx.x / (C * T);
  const int n = cu[b + 1] - cu[b];
  const size_t off = ((size_t)t * C + c) * Rp + pcol[b];
  const float4* lg = reinterpret_cast<const float4*>(logits + off);
  const float4* sc = reinterpret_cast<const float4*>(scores + off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float M = -INFINITY, Z = 0.f, Yv = 0.f;                     // running row statistics (meaningful in thread 0)
  for (int base = 0; base < n; base += 4 * ROW_THREADS * ROWC_V4) {
    float4 v[ROWC_V4], w[ROWC_V4];
#pragma unroll
    for (int k = 0; k < ROWC_V4; ++k) {
      const int i4 = base / 4 + threadIdx.x + ROW_THREADS * k;
      if (4 * i4 < n) { v[k] = __ldg(lg + i4); w[k] = __ldg(sc + i4); }
      else { v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); w[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < ROWC_V4; ++k) {
      const int col = base + 4 * (threadIdx.x + ROW_THREADS * k);
      if (col + 1 >= n) { v[k].y = -INFINITY; w[k].y = 0.f; }
      if (col + 2 >= n) { v[k].z = -INFINITY; w[k].z = 0.f; }
      if (col + 3 >= n) { v[k].w = -INFINITY; w[k].w = 0.f; }
      m = fmaxf(fmaxf(m, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
    m = warp_max(m);
    if (lane == 0) red[0][warp] = m;
    __syncthreads();
    m = red[0][0];
#pragma unroll
    for (int q = 1; q < ROW_THREADS / 32; ++q) m = fmaxf(m, red[0][q]);
    float z = 0.f, y = 0.f;
#pragma unroll
    for (int k = 0; k < ROWC_V4; ++k) {
      const float e0 = fast_exp(v[k].x - m), e1 = fast_exp(v[k].y - m);
      const float e2 = fast_exp(v[k].z - m), e3 = fast_exp(v[k].w - m);
      z += (e0 + e1) + (e2 + e3);
      y = fmaf(e0, w[k].x, y); y = fmaf(e1, w[k].y, y); y = fmaf(e2, w[k].z, y); y = fmaf(e3, w[k].w, y);
    }
    z = warp_sum(z); y = warp_sum(y);
    if (lane == 0) { red[1][warp] = z; red[2][warp] = y; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float zz = 0.f, yy = 0.f;
#pragma unroll
      for (int q = 0; q < ROW_THREADS / 32; ++q) { zz += red[1][q]; yy += red[2][q]; }
      const float Mn = fmaxf(M, m);
      const float fa = fast_exp(M - Mn), fb = fast_exp(m - Mn);             // exp(-inf) = 0 on the first chunk
      Z = Z * fa + zz * fb;
      Yv = Yv * fa + yy * fb;
      M = Mn;
    }
    __syncthreads();                                          // red[] is reused by the next chunk
  }
  if (threadIdx.x == 0) {
    const float inv = 1.0f / Z;
    rowstat[((size_t)c * n_bags + b) * T + t] = make_float2(M, inv);
    Y[((size_t)b * T + t) * C + c] = Yv * inv;
  }
}

// ================================================================================== generic path, launch 2: columns
constexpr int COL_LANES = 32;
constexpr int COL_VEC = 4;        // patches per lane: one 16-byte load per sample
constexpr int COL_COLS = COL_LANES * COL_VEC;   // = TILE_ROWS: one CTA works on one 128-patch tile of one bag
constexpr int COL_TGROUPS = 8;    // the CTA's samples are strided over 8 warps, then merged in a fixed order
constexpr int COL_THREADS = COL_LANES * COL_TGROUPS;
constexpr int COL_UNROLL = 4;     // samples a warp has in flight (independent 512-byte loads)
constexpr int COL_MAX_SPLIT = 16;
static_assert(COL_COLS == TILE_ROWS, "one column CTA per projection tile");

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

// grid (n_tiles + bag_blocks, C, wsplit).
// blockIdx.x < n_tiles: CTA = (tile of <= 128 patches of one bag, head c, sample group z): samples
//   [T z / wsplit, T (z+1) / wsplit).  Warp g takes the group's samples g, g+8, ...: every lane reads 4 adjacent patches
//   of a sample with one 16-byte load (512 contiguous bytes per warp and sample), four samples in flight per warp, no
//   shared-memory staging; A = exp(l - max) / sum and the Welford update run as packed fp32x2 operations (two patches
//   per instruction).  The 8 partial (count, mean, M2) per patch are merged with Chan's formula in warp order.
//   wsplit > 1 (few tiles, many samples: one large bag): every CTA writes its partial to the workspace and the
//   LAST CTA of the (tile, head) to arrive (one integer atomic per CTA) merges the groups in the order 0..wsplit-1,
//   so the result does not depend on the arrival order.  Optionally stores A.
// blockIdx.x >= n_tiles (y = z = 0): one warp per bag: mean / M2 over t of softmax_c(Y[bag][t][:])
template <bool HAS_A>
__global__ void __launch_bounds__(COL_THREADS)
welford_cols_kernel(const float* __restrict__ logits, const float2* __restrict__ rowstat,
                    const TileDesc* __restrict__ tiles, const float* __restrict__ Y,
                    int n_bags, int T, int C, int R, int Rp, int n_tiles, int wsplit,
                    float2* __restrict__ wpart, int* __restrict__ wcount,
                    float* __restrict__ A, float* __restrict__ attn_mean, float* __restrict__ attn_m2,
                    float* __restrict__ prob_mean, float* __restrict__ prob_m2) {
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  grid_dep_sync();                                           // rowstat / Y of the row kernel
  if ((int)blockIdx.x >= n_tiles) {
    if (blockIdx.y != 0 || blockIdx.z != 0 || prob_mean == nullptr) return;
    const int b = ((int)blockIdx.x - n_tiles) * COL_TGROUPS + grp;
    if (b >= n_bags) return;
    prob_stats_warp(Y + (size_t)b * T * C, T, C, lane, prob_mean + b * C, prob_m2 ? prob_m2 + b * C : nullptr);
    return;
  }
  __shared__ float s_mean[COL_TGROUPS][COL_COLS], s_m2[COL_TGROUPS][COL_COLS];
  __shared__ int s_cnt[COL_TGROUPS];
  __shared__ int s_last;
  const int nrows = tiles[blockIdx.x].nrows, pcol0 = tiles[blockIdx.x].pcol0, row0 = tiles[blockIdx.x].row0;
  const int bag = tiles[blockIdx.x].bag;
  const int c = blockIdx.y, z = blockIdx.z;
  const int t_lo = (int)((long long)T * z / wsplit), t_hi = (int)((long long)T * (z + 1) / wsplit);
  // (the last 16-byte load of the tile may reach into the bag's plane padding: loaded, its results never stored)
  const float* plane = logits + (size_t)c * Rp + pcol0 + lane * COL_VEC;
  const size_t tstride = (size_t)C * Rp;
  const float2* rs_row = rowstat + ((size_t)c * n_bags + bag) * T;
  uint64_t mean01 = 0ull, mean23 = 0ull, q01 = 0ull, q23 = 0ull;
  int cnt = 0;
  constexpr float LOG2E = 1.4426950408889634f;
  if (lane * COL_VEC < nrows) {
    for (int t = t_lo + grp; t < t_hi; t += COL_TGROUPS * COL_UNROLL) {
      float4 l4[COL_UNROLL];
      float2 rs[COL_UNROLL];
#pragma unroll
      for (int u = 0; u < COL_UNROLL; ++u) {
        const int tu = t + u * COL_TGROUPS;
        if (tu < t_hi) {
          l4[u] = __ldg(reinterpret_cast<const float4*>(plane + (size_t)tu * tstride));
          rs[u] = __ldg(rs_row + tu);
        }
      }
#pragma unroll
      for (int u = 0; u < COL_UNROLL; ++u) {
        const int tu = t + u * COL_TGROUPS;
        if (tu < t_hi) {
          // a = 2^(l log2e - max log2e) / sum, two patches per instruction
          const float nm = -rs[u].x * LOG2E;
          const uint64_t k2 = f2_pack(LOG2E, LOG2E), nm2 = f2_pack(nm, nm), inv2 = f2_pack(rs[u].y, rs[u].y);
          float x0, x1, x2, x3;
          f2_unpack(f2_fma(f2_pack(l4[u].x, l4[u].y), k2, nm2), x0, x1);
          f2_unpack(f2_fma(f2_pack(l4[u].z, l4[u].w), k2, nm2), x2, x3);
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
            float* dst = A + ((size_t)tu * C + c) * R + row0 + lane * COL_VEC;
            const int left = nrows - lane * COL_VEC;
            dst[0] = a0;
            if (left > 1) dst[1] = a1;
            if (left > 2) dst[2] = a2;
            if (left > 3) dst[3] = a3;
          }
          ++cnt;
          const float ic = fast_rcp((float)cnt);
          const uint64_t ic2 = f2_pack(ic, ic);
          const uint64_t d01 = f2_sub(a01, mean01), d23 = f2_sub(a23, mean23);
          mean01 = f2_fma(d01, ic2, mean01);
          mean23 = f2_fma(d23, ic2, mean23);
          q01 = f2_fma(d01, f2_sub(a01, mean01), q01);
          q23 = f2_fma(d23, f2_sub(a23, mean23), q23);
        }
      }
    }
  }
  {
    float m0, m1, m2_, m3, v0, v1, v2, v3;
    f2_unpack(mean01, m0, m1); f2_unpack(mean23, m2_, m3);
    f2_unpack(q01, v0, v1); f2_unpack(q23, v2, v3);
    *reinterpret_cast<float4*>(&s_mean[grp][lane * COL_VEC]) = make_float4(m0, m1, m2_, m3);
    *reinterpret_cast<float4*>(&s_m2[grp][lane * COL_VEC]) = make_float4(v0, v1, v2, v3);
    if (lane == 0) s_cnt[grp] = cnt;
  }
  __syncthreads();
  const int col = threadIdx.x;
  float mu = 0.f, q = 0.f, n_a = 0.f;
  if (col < nrows) {
#pragma unroll
    for (int k = 0; k < COL_TGROUPS; ++k) wf_merge(n_a, mu, q, (float)s_cnt[k], s_mean[k][col], s_m2[k][col]);
  }
  if (wsplit == 1) {
    if (col < nrows) {
      if (attn_mean) attn_mean[(size_t)c * R + row0 + col] = mu;
      if (attn_m2) attn_m2[(size_t)c * R + row0 + col] = q;
    }
    return;
  }
  if (col < nrows) wpart[((size_t)z * C + c) * Rp + pcol0 + col] = make_float2(mu, q);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&wcount[blockIdx.x * C + c], 1) == wsplit - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (col < nrows) {
    mu = 0.f; q = 0.f; n_a = 0.f;
    for (int k = 0; k < wsplit; ++k) {
      const float2 pk = __ldcg(&wpart[((size_t)k * C + c) * Rp + pcol0 + col]);
      const int nk = (int)((long long)T * (k + 1) / wsplit) - (int)((long long)T * k / wsplit);
      wf_merge(n_a, mu, q, (float)nk, pk.x, pk.y);
    }
    if (attn_mean) attn_mean[(size_t)c * R + row0 + col] = mu;
    if (attn_m2) attn_m2[(size_t)c * R + row0 + col] = q;
  }
}

int welford_split(int n_tiles, int C, int T) {
  // enough CTAs for ~4 per SM when there are few tiles (one large bag), at least 16 samples per group
  const long long ctas = (long long)n_tiles * C;
  long long s = (592 + ctas - 1) / ctas;
  if (s > COL_MAX_SPLIT) s = COL_MAX_SPLIT;
  if (s > T / 16) s = T / 16;
  return s < 1 ? 1 : (int)s;
}

// ================================================================================== one-launch path
// One cluster of 8 CTAs per bag.  CTA `rank` owns the patches [rank W, (rank+1) W), W = ceil(n/8) rounded up to 4:
//   stage   its [T*C][W] slabs of both planes into shared memory with 16-byte cp.async (everything in flight at once);
//   phase 1 one warp per (t, c) row (4 rows interleaved): max, sum exp, sum exp * score over the slab -> part[row];
//   phase 2 cluster barrier, then every CTA combines the 8 partials of every row through distributed shared
//           memory into the row's (max, 1 / sum); rank 0 writes Y and keeps it for the probability statistics;
//   phase 3 thread per (sample group, head, patch): Welford over t of exp(l - max) / sum from the resident slab,
//           Chan merge of the FR_TG sample groups, optional A store; rank 0 / warp 0: statistics of softmax_c(Y).
// Both planes are read exactly once (from L2 when the call follows the projection of a single bag).
constexpr int FR_CL = 8;
constexpr int FR_THREADS = 512;
constexpr int FR_TG = 2;          // sample groups of phase 3 (t = g, g + 2, ...)
constexpr int FR_RPI = 8;         // rows a warp works on at a time in phase 1 (independent shuffle chains)
constexpr size_t FR_SMEM_MAX = 220 * 1024;

__host__ __device__ inline int fr_slab_cols(int n) { return ((n + FR_CL - 1) / FR_CL + 3) & ~3; }

size_t fused_reduce_smem_bytes(int T, int C, int max_n) {
  const size_t rows = (size_t)T * C, Wp = (size_t)fr_slab_cols(max_n);
  const size_t rows2 = (rows + 1) & ~(size_t)1;
  const size_t bytes = 2 * rows * Wp * 4 + rows * 16 + rows * 8 + rows2 * 4 + (size_t)FR_TG * C * Wp * 8;
  return bytes <= FR_SMEM_MAX ? bytes : 0;
}

__device__ __forceinline__ uint32_t fr_cluster_rank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void fr_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 fr_ld_remote4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

template <bool HAS_A>
__global__ void __cluster_dims__(FR_CL, 1, 1) __launch_bounds__(FR_THREADS, 1)
fused_bag_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                        const int32_t* __restrict__ cu, const int32_t* __restrict__ pcol, int T, int C, int R,
                        int Rp, int Wp, float* __restrict__ Y, float* __restrict__ A,
                        float* __restrict__ attn_mean, float* __restrict__ attn_m2,
                        float* __restrict__ prob_mean, float* __restrict__ prob_m2) {
  extern __shared__ __align__(16) uint8_t fr_smem[];
  const int rows = T * C, rows2 = (rows + 1) & ~1;
  float* slabL = reinterpret_cast<float*>(fr_smem);
  float* slabS = slabL + (size_t)rows * Wp;
  float4* part = reinterpret_cast<float4*>(slabS + (size_t)rows * Wp);
  float2* stat = reinterpret_cast<float2*>(part + rows);
  float* yv = reinterpret_cast<float*>(stat + rows);
  float2* p3 = reinterpret_cast<float2*>(yv + rows2);         // [FR_TG][C][Wp]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)fr_cluster_rank();
  const int b = blockIdx.x / FR_CL;
  const int r0 = cu[b], n = cu[b + 1] - r0;
  const int W = fr_slab_cols(n);
  const int col0 = rank * W;
  const int ncols = max(0, min(W, n - col0));
  const int nv4 = (ncols + 3) >> 2;
  grid_dep_sync();

  // ---- stage both slabs (the last 16-byte piece of the bag may reach into its plane padding: copied, masked below)
  {
    const size_t goff = (size_t)pcol[b] + col0;
    const uint32_t sL = (uint32_t)__cvta_generic_to_shared(slabL), sS = (uint32_t)__cvta_generic_to_shared(slabS);
    for (int row = warp; row < rows; row += FR_THREADS / 32)
      for (int s4 = lane; s4 < nv4; s4 += 32) {
        const size_t g = (size_t)row * Rp + goff + 4 * s4;
        const uint32_t so = (uint32_t)(row * Wp + 4 * s4) * 4u;
        cp_async16(sL + so, logits + g);
        cp_async16(sS + so, scores + g);
      }
    cp_async_wait_all();
    __syncthreads();
  }

  // ---- phase 1: per-row partials over this CTA's slab
  for (int rb = warp * FR_RPI; rb < rows; rb += (FR_THREADS / 32) * FR_RPI) {
    float m[FR_RPI], z[FR_RPI], y[FR_RPI];
#pragma unroll
    for (int j = 0; j < FR_RPI; ++j) {
      m[j] = -INFINITY;
      const int row = rb + j;
      if (row < rows)
        for (int s4 = lane; s4 < nv4; s4 += 32) {
          const float4 l = *reinterpret_cast<const float4*>(slabL + (size_t)row * Wp + 4 * s4);
          const int cc = 4 * s4;
          m[j] = fmaxf(m[j], l.x);
          if (cc + 1 < ncols) m[j] = fmaxf(m[j], l.y);
          if (cc + 2 < ncols) m[j] = fmaxf(m[j], l.z);
          if (cc + 3 < ncols) m[j] = fmaxf(m[j], l.w);
        }
    }
#pragma unroll
    for (int j = 0; j < FR_RPI; ++j) m[j] = warp_max(m[j]);
#pragma unroll
    for (int j = 0; j < FR_RPI; ++j) {
      z[j] = 0.f; y[j] = 0.f;
      const int row = rb + j;
      if (row < rows)
        for (int s4 = lane; s4 < nv4; s4 += 32) {
          const float4 l = *reinterpret_cast<const float4*>(slabL + (size_t)row * Wp + 4 * s4);
          const float4 s = *reinterpret_cast<const float4*>(slabS + (size_t)row * Wp + 4 * s4);
          const int cc = 4 * s4;
          const float e0 = fast_exp(l.x - m[j]);
          const float e1 = cc + 1 < ncols ? fast_exp(l.y - m[j]) : 0.f;
          const float e2 = cc + 2 < ncols ? fast_exp(l.z - m[j]) : 0.f;
          const float e3 = cc + 3 < ncols ? fast_exp(l.w - m[j]) : 0.f;
          z[j] += (e0 + e1) + (e2 + e3);
          y[j] = fmaf(e0, s.x, y[j]);
          if (cc + 1 < ncols) y[j] = fmaf(e1, s.y, y[j]);
          if (cc + 2 < ncols) y[j] = fmaf(e2, s.z, y[j]);
          if (cc + 3 < ncols) y[j] = fmaf(e3, s.w, y[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < FR_RPI; ++j) { z[j] = warp_sum(z[j]); y[j] = warp_sum(y[j]); }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < FR_RPI; ++j)
        if (rb + j < rows) part[rb + j] = make_float4(m[j], z[j], y[j], 0.f);     // (-inf, 0, 0) for an empty slab
    }
  }
  fr_cluster_sync();                                          // part[] of all 8 CTAs is complete and visible

  // ---- phase 2: combine the 8 slabs of every row (every CTA does this redundantly: 8 remote 16-byte loads per row)
  {
    const uint32_t part_s = (uint32_t)__cvta_generic_to_shared(part);
    for (int row = threadIdx.x; row < rows; row += FR_THREADS) {
      float4 pk[FR_CL];
#pragma unroll
      for (int r = 0; r < FR_CL; ++r) pk[r] = fr_ld_remote4(part_s + (uint32_t)row * 16u, (uint32_t)r);
      float M = pk[0].x;
#pragma unroll
      for (int r = 1; r < FR_CL; ++r) M = fmaxf(M, pk[r].x);
      float Z = 0.f, Ys = 0.f;
#pragma unroll
      for (int r = 0; r < FR_CL; ++r) {
        const float f = fast_exp(pk[r].x - M);                // exp(-inf) = 0 for an empty slab
        Z = fmaf(pk[r].y, f, Z);
        Ys = fmaf(pk[r].z, f, Ys);
      }
      const float inv = 1.0f / Z;
      stat[row] = make_float2(M, inv);
      yv[row] = Ys * inv;
      if (rank == 0) Y[(size_t)b * rows + row] = Ys * inv;    // [n_bags][T][C], row = t*C + c
    }
    __syncthreads();
  }

  // ---- phase 3: attention values of the slab, Welford over the samples (4 samples in flight per thread: the
  // loads, exponentials and reciprocal counts are independent, only the two-instruction Welford chain is serial)
  const int items = ncols * C;
  for (int it = threadIdx.x; it < items * FR_TG; it += FR_THREADS) {
    const int col = it % ncols, rest = it / ncols;
    const int c = rest % C, tg = rest / C;
    float mean = 0.f, m2 = 0.f;
    int cnt = 0;
    const float* lp = slabL + (size_t)c * Wp + col;
    const size_t rstride = (size_t)C * Wp;
    for (int t = tg; t < T; t += 4 * FR_TG) {
      float a[4], ic[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tu = t + u * FR_TG;
        if (tu < T) {
          const float2 st = stat[tu * C + c];
          a[u] = fast_exp(lp[(size_t)tu * rstride] - st.x) * st.y;
          ic[u] = fast_rcp((float)(cnt + u + 1));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tu = t + u * FR_TG;
        if (tu < T) {
          if constexpr (HAS_A) A[(size_t)(tu * C + c) * R + r0 + col0 + col] = a[u];
          wf_push(mean, m2, a[u], ic[u]);
        }
      }
      cnt += 4;
    }
    p3[((size_t)tg * C + c) * Wp + col] = make_float2(mean, m2);
  }
  __syncthreads();
  for (int it = threadIdx.x; it < items; it += FR_THREADS) {
    const int col = it % ncols, c = it / ncols;
    float mu = 0.f, q = 0.f, n_a = 0.f;
#pragma unroll
    for (int tg = 0; tg < FR_TG; ++tg) {
      const float2 pk = p3[((size_t)tg * C + c) * Wp + col];
      wf_merge(n_a, mu, q, (float)((T - tg + FR_TG - 1) / FR_TG), pk.x, pk.y);
    }
    const size_t o = (size_t)c * R + r0 + col0 + col;
    if (attn_mean) attn_mean[o] = mu;
    if (attn_m2) attn_m2[o] = q;
  }
  // ---- statistics of softmax_c(Y) over the samples: rank 0, one thread per sample, two block reductions
  // (mean first, then the squared deviations: the two-pass form of the column kernel's prob_stats_warp)
  if (rank == 0 && prob_mean != nullptr) {
    __shared__ float s_red[2][FR_THREADS / 32][MAXC];
    float s[MAXC] = {0.f, 0.f, 0.f, 0.f};
    for (int t = threadIdx.x; t < T; t += FR_THREADS) {
      float pt[MAXC];
      softmax_classes(yv + t * C, C, pt);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) s[c] += pt[c];
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) s[c] = warp_sum(s[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) s_red[0][warp][c] = s[c];
    }
    __syncthreads();
    float mean[MAXC], q[MAXC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < FR_THREADS / 32; ++w) tot += s_red[0][w][c];
      mean[c] = tot / (float)T;
    }
    for (int t = threadIdx.x; t < T; t += FR_THREADS) {
      float pt[MAXC];
      softmax_classes(yv + t * C, C, pt);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) { const float dlt = pt[c] - mean[c]; q[c] = fmaf(dlt, dlt, q[c]); }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) q[c] = warp_sum(q[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) s_red[1][warp][c] = q[c];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < FR_THREADS / 32; ++w) tot += s_red[1][w][threadIdx.x];
      float mc = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) mc = c == (int)threadIdx.x ? mean[c] : mc;
      prob_mean[b * C + threadIdx.x] = mc;
      if (prob_m2) prob_m2[b * C + threadIdx.x] = tot;
    }
  }
  fr_cluster_sync();                                          // no CTA leaves while a peer may still read its part[]
}

// dynamic shared memory opt-in is a per-device function attribute: set it once per (kernel, device)
static cudaError_t fused_attr(const void* fn) {
  static std::mutex mu;
  static bool done[2][64] = {};
  static const void* fns[2] = {nullptr, nullptr};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  int slot = fns[0] == fn ? 0 : (fns[1] == fn ? 1 : (fns[0] == nullptr ? 0 : 1));
  fns[slot] = fn;
  if (dev < 0 || dev >= 64 || !done[slot][dev]) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM_MAX);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) done[slot][dev] = true;
  }
  return cudaSuccess;
}

cudaError_t launch_reduce(const Plan& p, const float* logits, const float* scores, uint8_t* workspace,
                          float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean,
                          float* attn_m2, int reduce_path, cudaStream_t st, int* launches) {
  const int32_t* cu = p.d_cu;
  const int32_t* pcol = p.d_pcol;
  int n_bags = p.n_bags, T = p.T, C = p.C, R = p.R, Rp = p.Rp;
  // ---- one launch: latency path (small batches whose slabs fit shared memory)
  const bool fused_fits = p.fused_smem != 0;
  if (reduce_path == 2 && !fused_fits) return cudaErrorInvalidConfiguration;
  if (reduce_path == 2 || (reduce_path == 0 && fused_fits && p.n_bags * FR_CL <= 4 * 148)) {
    int Wp = fr_slab_cols(p.max_n);
    const void* fn = A != nullptr ? (const void*)fused_bag_reduce_kernel<true> : (const void*)fused_bag_reduce_kernel<false>;
    cudaError_t e = fused_attr(fn);
    if (e != cudaSuccess) return e;
    PdlLaunch L(dim3((unsigned)(p.n_bags * FR_CL)), dim3(FR_THREADS), p.fused_smem, st);
    e = A != nullptr
        ? cudaLaunchKernelEx(&L.cfg, fused_bag_reduce_kernel<true>, logits, scores, cu, pcol, T, C, R, Rp, Wp, Y, A,
                             attn_mean, attn_m2, prob_mean, prob_m2)
        : cudaLaunchKernelEx(&L.cfg, fused_bag_reduce_kernel<false>, logits, scores, cu, pcol, T, C, R, Rp, Wp, Y, A,
                             attn_mean, attn_m2, prob_mean, prob_m2);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaGetLastError();
  }
  // ---- two launches: rows, then columns
  float2* rowstat = reinterpret_cast<float2*>(workspace + p.off_rowstat);
  float2* wpart = reinterpret_cast<float2*>(workspace + p.off_wpart);
  int* wcount = reinterpret_cast<int*>(workspace + p.off_wcount);
  int n_wcount = p.wsplit > 1 ? p.n_tiles * p.C : 0;
  {
    const long long rows = (long long)p.n_bags * p.T * p.C;
    const unsigned warp_grid = (unsigned)((rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32));
    cudaError_t e;
    // one warp per row when there are enough rows to fill the GPU that way and the rows are not so long that a
    // CTA per row streams them better
    if (p.max_n <= 8192 && rows >= 2048) {
      PdlLaunch L(dim3(warp_grid), dim3(ROW_THREADS), 0, st);
      e = cudaLaunchKernelEx(&L.cfg, softmax_rows_warp_kernel, logits, scores, cu, pcol, n_bags, T, C, Rp, rowstat, Y, wcount, n_wcount);
    } else {
      PdlLaunch L(dim3((unsigned)rows), dim3(ROW_THREADS), 0, st);
      e = cudaLaunchKernelEx(&L.cfg, softmax_rows_cta_kernel, logits, scores, cu, pcol, n_bags, T, C, Rp, rowstat, Y, wcount, n_wcount);
    }
    if (e != cudaSuccess) return e;
  }
  if (launches) ++*launches;
  const int bag_blocks = (p.n_bags + COL_TGROUPS - 1) / COL_TGROUPS;
  {
    PdlLaunch L(dim3(p.n_tiles + bag_blocks, p.C, p.wsplit), dim3(COL_THREADS), 0, st);
    const float2* rs = rowstat;
    const TileDesc* tiles = p.d_tiles;
    const float* Yc = Y;
    int n_tiles = p.n_tiles, wsplit = p.wsplit;
    cudaError_t e = A != nullptr
        ? cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<true>, logits, rs, tiles, Yc, n_bags, T, C, R, Rp, n_tiles, wsplit,
                             wpart, wcount, A, attn_mean, attn_m2, prob_mean, prob_m2)
        : cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<false>, logits, rs, tiles, Yc, n_bags, T, C, R, Rp, n_tiles, wsplit,
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

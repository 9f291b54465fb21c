// reduce.cu — the reductions behind the projection: per-head softmax over the variable-length
// patch axis (model.py:305), the attention-weighted classifier contraction (model.py:308-316,
// as sum_n A[t,c,n]*score[t,c,n]), and the Welford mean / M2 over the T MC samples of the class
// probabilities (infer.py:195, net_utils.py:207-208) and of the attention (infer.py:212-219).
// Warp-shuffle kernels, coalesced along the patch axis, no atomics (run-to-run deterministic).
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

constexpr int ROW_THREADS = 256;

// one CTA per (bag, t, c) row: rowstat = (max, 1/sum exp), Y[bag][t][c] = sum_n softmax_n * score
__global__ void __launch_bounds__(ROW_THREADS)
softmax_rows_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                    const int32_t* __restrict__ cu, int n_bags, int T, int C, int Rp,
                    float2* __restrict__ rowstat, float* __restrict__ Y) {
  __shared__ float red[2][ROW_THREADS / 32];
  asm volatile("griddepcontrol.wait;" ::: "memory");         // logits / scores of the projection kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int c = blockIdx.x % C;
  const int t = (blockIdx.x / C) % T;
  const int b = blockIdx.x / (C * T);
  const int r0 = cu[b], n = cu[b + 1] - r0;
  const float* lg = logits + ((size_t)t * C + c) * Rp + r0;
  const float* sc = scores + ((size_t)t * C + c) * Rp + r0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += ROW_THREADS) m = fmaxf(m, lg[i]);
  m = warp_max(m);
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int w = 1; w < ROW_THREADS / 32; ++w) m = fmaxf(m, red[0][w]);
  __syncthreads();

  float z = 0.f, y = 0.f;
  for (int i = threadIdx.x; i < n; i += ROW_THREADS) {
    const float e = fast_exp(lg[i] - m);
    z += e;
    y = fmaf(e, sc[i], y);
  }
  z = warp_sum(z); y = warp_sum(y);
  if (lane == 0) { red[0][warp] = z; red[1][warp] = y; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float zz = 0.f, yy = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_THREADS / 32; ++w) { zz += red[0][w]; yy += red[1][w]; }
    const float inv = 1.0f / zz;
    rowstat[((size_t)c * n_bags + b) * T + t] = make_float2(m, inv);   // [C][n_bags][T]: a row's samples are contiguous
    Y[((size_t)b * T + t) * C + c] = yy * inv;
  }
}

// Short rows (every bag <= 32 * VPL patches): one WARP per (bag, t, c) row, eight rows per CTA.  The logits of
// the row stay in registers between the max and the exp / sum pass (each plane is read from DRAM exactly
// once), no shared memory, no __syncthreads; the fixed shuffle order keeps the result deterministic.
template <int VPL>
__global__ void __launch_bounds__(ROW_THREADS)
softmax_rows_warp_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                         const int32_t* __restrict__ cu, int n_bags, int T, int C, int Rp,
                         float2* __restrict__ rowstat, float* __restrict__ Y) {
  asm volatile("griddepcontrol.wait;" ::: "memory");         // logits / scores of the projection kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= (long long)n_bags * T * C) return;
  const int c = (int)(row % C);
  const int t = (int)((row / C) % T);
  const int b = (int)(row / ((long long)C * T));
  const int r0 = cu[b], n = cu[b + 1] - r0;
  const float* lg = logits + ((size_t)t * C + c) * Rp + r0;
  const float* sc = scores + ((size_t)t * C + c) * Rp + r0;
  float v[VPL], w[VPL];                                      // both planes of the row in flight at once
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < n ? __ldg(lg + i) : -INFINITY;
    w[k] = i < n ? __ldg(sc + i) : 0.f;
  }
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < VPL; ++k) m = fmaxf(m, v[k]);
  m = warp_max(m);
  float z = 0.f, y = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const float e = fast_exp(v[k] - m);                      // exp(-inf) = 0 for the padding
    z += e;
    y = fmaf(e, w[k], y);
  }
  z = warp_sum(z); y = warp_sum(y);
  if (lane == 0) {
    const float inv = 1.0f / z;
    rowstat[((size_t)c * n_bags + b) * T + t] = make_float2(m, inv);   // [C][n_bags][T]: a row's samples are contiguous
    Y[((size_t)b * T + t) * C + c] = y * inv;
  }
}

constexpr int COL_LANES = 32;
constexpr int COL_VEC = 4;        // packed rows (patches) per lane: one 16-byte load per sample
constexpr int COL_COLS = COL_LANES * COL_VEC;   // packed rows per CTA
constexpr int COL_TGROUPS = 8;    // MC samples are strided over 8 warps, then merged in a fixed order
constexpr int COL_THREADS = COL_LANES * COL_TGROUPS;

constexpr int COL_CHUNK_T = 64;   // samples staged in shared memory at a time (32 KB)

// grid.x < col_blocks: CTA = 128 packed rows x one head.  The CTA's slab of the logit planes ([T][128] floats,
//   512 contiguous bytes per sample) is staged through shared memory with cp.async in chunks of 64 samples:
//   every thread has 8 independent 16-byte copies in flight without holding registers, several CTAs per SM
//   overlap their copy and compute phases.  Warp g then runs Welford over the samples t = g, g+8, ... of
//   A[t,c,row], each lane on 4 adjacent rows, and the 8 partial (count, mean, M2) per row are merged with
//   Chan's formula in warp order 0..7 (deterministic).  Optionally stores A.
// grid.x >= col_blocks: one warp per bag: mean / M2 over t of softmax_c(Y[bag][t][:])
template <bool HAS_A>
__global__ void __launch_bounds__(COL_THREADS)
welford_cols_kernel(const float* __restrict__ logits, const float2* __restrict__ rowstat,
                    const int32_t* __restrict__ row2bag, const float* __restrict__ Y,
                    int n_bags, int T, int C, int R, int Rp, int col_blocks,
                    float* __restrict__ A, float* __restrict__ attn_mean, float* __restrict__ attn_m2,
                    float* __restrict__ prob_mean, float* __restrict__ prob_m2) {
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.wait;" ::: "memory");         // rowstat / Y of softmax_rows_kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((int)blockIdx.x < col_blocks) {
    __shared__ __align__(16) float s_lg[COL_CHUNK_T][COL_COLS];
    __shared__ float s_mean[COL_TGROUPS][COL_COLS], s_m2[COL_TGROUPS][COL_COLS];
    const int c = blockIdx.y;
    const int col0 = blockIdx.x * COL_COLS;
    const int g0 = col0 + lane * COL_VEC;                    // multiple of 4; the planes have stride Rp (multiple of 32)
    float mean[COL_VEC] = {0.f, 0.f, 0.f, 0.f}, m2[COL_VEC] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    int b[COL_VEC];
#pragma unroll
    for (int k = 0; k < COL_VEC; ++k) b[k] = row2bag[min(g0 + k, R - 1)];
    const bool warp_mixed = __any_sync(0xffffffffu, b[0] != b[COL_VEC - 1]);
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(&s_lg[0][0]);
    for (int t0 = 0; t0 < T; t0 += COL_CHUNK_T) {
      const int nt = min(COL_CHUNK_T, T - t0);
      // 16-byte pieces of the [nt][128] slab; columns >= Rp do not exist (columns in [R, Rp) are plane padding:
      // copied, never used)
      for (int p = threadIdx.x; p < nt * (COL_COLS / 4); p += COL_THREADS) {
        const int row = p / (COL_COLS / 4), seg = p % (COL_COLS / 4);
        const int gcol = col0 + seg * 4;
        if (gcol < Rp) {
          const float* src = logits + ((size_t)(t0 + row) * C + c) * Rp + gcol;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s_base + (uint32_t)(row * COL_COLS + seg * 4) * 4u), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (g0 < R) {
        const float2* rs_row = rowstat + ((size_t)c * n_bags + b[0]) * T + t0;
#pragma unroll 4
        for (int r = grp; r < nt; r += COL_TGROUPS) {
          const float4 l4 = *reinterpret_cast<const float4*>(&s_lg[r][lane * COL_VEC]);
          const float lg[COL_VEC] = {l4.x, l4.y, l4.z, l4.w};
          float2 rs[COL_VEC];
          rs[0] = __ldg(rs_row + r);
#pragma unroll
          for (int k = 1; k < COL_VEC; ++k) rs[k] = rs[0];
          if (warp_mixed) {                                  // a bag boundary inside this warp's 128 rows (warp-uniform branch)
#pragma unroll
            for (int k = 1; k < COL_VEC; ++k)
              if (b[k] != b[0]) rs[k] = __ldg(rowstat + ((size_t)c * n_bags + b[k]) * T + t0 + r);
          }
          ++cnt;
          const float inv_cnt = fast_rcp((float)cnt);
#pragma unroll
          for (int k = 0; k < COL_VEC; ++k) {
            const float a = fast_exp(lg[k] - rs[k].x) * rs[k].y;
            if constexpr (HAS_A) {
              if (g0 + k < R) A[((size_t)(t0 + r) * C + c) * R + g0 + k] = a;
            }
            const float dlt = a - mean[k];
            mean[k] += dlt * inv_cnt;
            m2[k] = fmaf(dlt, a - mean[k], m2[k]);
          }
        }
      }
      __syncthreads();                                       // the next chunk overwrites s_lg
    }
#pragma unroll
    for (int k = 0; k < COL_VEC; ++k) { s_mean[grp][lane * COL_VEC + k] = mean[k]; s_m2[grp][lane * COL_VEC + k] = m2[k]; }
    __syncthreads();
    const int col = threadIdx.x, g = blockIdx.x * COL_COLS + col;
    if (col < COL_COLS && g < R) {
      float mu = s_mean[0][col], q = s_m2[0][col];
      float n_a = (float)((T + COL_TGROUPS - 1) / COL_TGROUPS);   // group 0 always has the most samples
#pragma unroll
      for (int k = 1; k < COL_TGROUPS; ++k) {
        const int nk = (T - k + COL_TGROUPS - 1) / COL_TGROUPS;   // samples of group k
        if (nk <= 0) continue;
        const float n_b = (float)nk, mb = s_mean[k][col], qb = s_m2[k][col];
        const float n_ab = n_a + n_b, dlt = mb - mu;
        mu += dlt * __fdividef(n_b, n_ab);
        q += qb + dlt * dlt * __fdividef(n_a * n_b, n_ab);
        n_a = n_ab;
      }
      if (attn_mean) attn_mean[(size_t)c * R + g] = mu;
      if (attn_m2) attn_m2[(size_t)c * R + g] = q;
    }
  } else {
    if (blockIdx.y != 0 || prob_mean == nullptr) return;
    const int b = ((int)blockIdx.x - col_blocks) * COL_TGROUPS + grp;
    if (b >= n_bags) return;
    const float* y = Y + (size_t)b * T * C;
    float s[MAXC] = {0.f, 0.f, 0.f, 0.f};
    for (int t = lane; t < T; t += 32) {
      float mx = -INFINITY, p[MAXC], z = 0.f;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, y[t * C + c]);
      for (int c = 0; c < C; ++c) { p[c] = __expf(y[t * C + c] - mx); z += p[c]; }
      for (int c = 0; c < C; ++c) s[c] += p[c] / z;
    }
    float mean[MAXC];
    for (int c = 0; c < C; ++c) mean[c] = warp_sum(s[c]) / (float)T;
    float q[MAXC] = {0.f, 0.f, 0.f, 0.f};
    for (int t = lane; t < T; t += 32) {
      float mx = -INFINITY, p[MAXC], z = 0.f;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, y[t * C + c]);
      for (int c = 0; c < C; ++c) { p[c] = __expf(y[t * C + c] - mx); z += p[c]; }
      for (int c = 0; c < C; ++c) { const float dlt = p[c] / z - mean[c]; q[c] = fmaf(dlt, dlt, q[c]); }
    }
    for (int c = 0; c < C; ++c) {
      const float qq = warp_sum(q[c]);
      if (lane == 0) { prob_mean[b * C + c] = mean[c]; if (prob_m2) prob_m2[b * C + c] = qq; }
    }
  }
}

cudaError_t launch_reduce(const Plan& p, const float* logits, const float* scores, float2* rowstat,
                          float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean,
                          float* attn_m2, cudaStream_t st, int* launches) {
  {
    const int32_t* cu = p.d_cu;
    int n_bags = p.n_bags, T = p.T, C = p.C, Rp = p.Rp;
    const long long rows = (long long)p.n_bags * p.T * p.C;
    cudaError_t e;
    if (p.max_n <= 32 * 32 && rows >= 4096) {        // enough short rows to fill the GPU with one warp per row
      PdlLaunch L(dim3((unsigned)((rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32))), dim3(ROW_THREADS), 0, st);
      e = cudaLaunchKernelEx(&L.cfg, softmax_rows_warp_kernel<32>, logits, scores, cu, n_bags, T, C, Rp, rowstat, Y);
    } else {
      PdlLaunch L(dim3((unsigned)rows), dim3(ROW_THREADS), 0, st);
      e = cudaLaunchKernelEx(&L.cfg, softmax_rows_kernel, logits, scores, cu, n_bags, T, C, Rp, rowstat, Y);
    }
    if (e != cudaSuccess) return e;
  }
  if (launches) ++*launches;
  const int col_blocks = (p.R + COL_COLS - 1) / COL_COLS;
  const int bag_blocks = (p.n_bags + COL_TGROUPS - 1) / COL_TGROUPS;
  {
    PdlLaunch L(dim3(col_blocks + bag_blocks, p.C), dim3(COL_THREADS), 0, st);
    const float2* rs = rowstat;
    const int32_t* r2b = p.d_row2bag;
    const float* Yc = Y;
    int n_bags = p.n_bags, T = p.T, C = p.C, R = p.R, Rp = p.Rp;
    cudaError_t e = A != nullptr
        ? cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<true>, logits, rs, r2b, Yc, n_bags, T, C, R, Rp, col_blocks,
                             A, attn_mean, attn_m2, prob_mean, prob_m2)
        : cudaLaunchKernelEx(&L.cfg, welford_cols_kernel<false>, logits, rs, r2b, Yc, n_bags, T, C, R, Rp, col_blocks,
                             A, attn_mean, attn_m2, prob_mean, prob_m2);
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

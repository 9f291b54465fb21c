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

constexpr int ROW_THREADS = 256;

// one CTA per (bag, t, c) row: rowstat = (max, 1/sum exp), Y[bag][t][c] = sum_n softmax_n * score
__global__ void __launch_bounds__(ROW_THREADS)
softmax_rows_kernel(const float* __restrict__ logits, const float* __restrict__ scores,
                    const int32_t* __restrict__ cu, int n_bags, int T, int C, int Rp,
                    float2* __restrict__ rowstat, float* __restrict__ Y) {
  __shared__ float red[2][ROW_THREADS / 32];
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
    const float e = __expf(lg[i] - m);
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
    rowstat[((size_t)t * C + c) * n_bags + b] = make_float2(m, inv);
    Y[((size_t)b * T + t) * C + c] = yy * inv;
  }
}

constexpr int COL_THREADS = 128;

// grid.x < col_blocks: one thread per (c, packed row): Welford over t of A[t,c,row] (+ optional A store)
// grid.x >= col_blocks: one warp per bag: mean / M2 over t of softmax_c(Y[bag][t][:])
__global__ void __launch_bounds__(COL_THREADS)
welford_cols_kernel(const float* __restrict__ logits, const float2* __restrict__ rowstat,
                    const int32_t* __restrict__ row2bag, const float* __restrict__ Y,
                    int n_bags, int T, int C, int R, int Rp, int col_blocks,
                    float* __restrict__ A, float* __restrict__ attn_mean, float* __restrict__ attn_m2,
                    float* __restrict__ prob_mean, float* __restrict__ prob_m2) {
  if ((int)blockIdx.x < col_blocks) {
    const int c = blockIdx.y;
    const int g = blockIdx.x * COL_THREADS + threadIdx.x;
    if (g >= R) return;
    const int b = row2bag[g];
    float mean = 0.f, m2 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float2 rs = rowstat[((size_t)t * C + c) * n_bags + b];
      const float a = __expf(logits[((size_t)t * C + c) * Rp + g] - rs.x) * rs.y;
      if (A) A[((size_t)t * C + c) * R + g] = a;
      const float dlt = a - mean;
      mean += __fdividef(dlt, (float)(t + 1));
      m2 = fmaf(dlt, a - mean, m2);
    }
    if (attn_mean) attn_mean[(size_t)c * R + g] = mean;
    if (attn_m2) attn_m2[(size_t)c * R + g] = m2;
  } else {
    if (blockIdx.y != 0 || prob_mean == nullptr) return;
    const int b = ((int)blockIdx.x - col_blocks) * (COL_THREADS / 32) + (threadIdx.x >> 5);
    if (b >= n_bags) return;
    const int lane = threadIdx.x & 31;
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
  softmax_rows_kernel<<<p.n_bags * p.T * p.C, ROW_THREADS, 0, st>>>(logits, scores, p.d_cu, p.n_bags, p.T, p.C,
                                                                    p.Rp, rowstat, Y);
  if (launches) ++*launches;
  const int col_blocks = (p.R + COL_THREADS - 1) / COL_THREADS;
  const int bag_blocks = (p.n_bags + COL_THREADS / 32 - 1) / (COL_THREADS / 32);
  welford_cols_kernel<<<dim3(col_blocks + bag_blocks, p.C), COL_THREADS, 0, st>>>(
      logits, rowstat, p.d_row2bag, Y, p.n_bags, p.T, p.C, p.R, p.Rp, col_blocks, A, attn_mean, attn_m2,
      prob_mean, prob_m2);
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

}  // namespace mcmil

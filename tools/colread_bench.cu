// colread_bench.cu — access-pattern microbenchmark for the column (Welford-over-samples) kernel.
//
// The column kernel walks a [T][C][Rp] fp32 plane along t for fixed columns.  This program reads the same plane with
// the same thread mapping but only sums what it loads, for several shapes of the per-warp access: K float4 per lane
// contiguous along the row (K x 512 bytes per warp and sample), U samples in flight per warp, TG sample groups per
// CTA (the other 8 / TG warps extend the CTA's contiguous span), SPLIT sample ranges per column block.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/colread_bench tools/colread_bench.cu
//   build/colread_bench T C Rp            (config 4: 1000 2 16384; config 2: 100 2 262144)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int K, int U>
__global__ void __launch_bounds__(256) colread(const float* __restrict__ plane, int T, int C, int Rp, int TG, int split,
                                               float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CW = 8 / TG;                         // column warps per CTA
  const int cw = warp % CW, grp = warp / CW;
  const int c = blockIdx.y, z = blockIdx.z;
  const int t_lo = (int)((long long)T * z / split), t_hi = (int)((long long)T * (z + 1) / split);
  const size_t col0 = ((size_t)blockIdx.x * CW + cw) * (K * 128);
  if (col0 >= (size_t)Rp) return;
  const float4* lp = reinterpret_cast<const float4*>(plane + (size_t)c * Rp + col0) + lane;
  const size_t tstride4 = (size_t)C * Rp / 4;
  float acc = 0.f;
  int t = t_lo + grp;
  lp += (size_t)t * tstride4;
  const size_t step = (size_t)TG * tstride4;
  for (; t + (U - 1) * TG < t_hi; t += U * TG) {
    float4 v[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) v[u][k] = __ldg(lp + u * step + k * 32);
    lp += U * step;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) acc += (v[u][k].x + v[u][k].y) + (v[u][k].z + v[u][k].w);
  }
  for (; t < t_hi; t += TG) {
#pragma unroll
    for (int k = 0; k < K; ++k) { const float4 v = __ldg(lp + k * 32); acc += (v.x + v.y) + (v.z + v.w); }
    lp += step;
  }
  if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) linread(const float4* __restrict__ p, size_t n4, float* __restrict__ out) {
  float acc = 0.f;
  size_t i = (size_t)blockIdx.x * 256 * 4 + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256 * 4;
  for (; i + 768 < n4; i += stride) {
    const float4 a = __ldg(p + i), b = __ldg(p + i + 256), c = __ldg(p + i + 512), d = __ldg(p + i + 768);
    acc += (a.x + b.y) + (c.z + d.w);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int K, int U>
static void run(const float* plane, int T, int C, int Rp, int TG, int split, float* out, size_t bytes, char* flush,
                size_t flush_bytes) {
  const int CW = 8 / TG;
  const int blocks = (Rp + CW * K * 128 - 1) / (CW * K * 128);
  dim3 grid(blocks, C, split);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(flush, it, flush_bytes);            // evict the plane from L2
    cudaEventRecord(e0);
    colread<K, U><<<grid, 256>>>(plane, T, C, Rp, TG, split, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  printf("K=%d U=%d TG=%d split=%2d ctas=%5d : %7.1f us  %6.2f TB/s\n", K, U, TG, split, blocks * C * split, best * 1e3,
         bytes / best / 1e9);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
}

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 1000, C = argc > 2 ? atoi(argv[2]) : 2, Rp = argc > 3 ? atoi(argv[3]) : 16384;
  const size_t n = (size_t)T * C * Rp, bytes = n * 4;
  float* plane; float* out; char* flush;
  const size_t flush_bytes = 512u << 20;
  cudaMalloc(&plane, bytes); cudaMalloc(&out, 4); cudaMalloc(&flush, flush_bytes);
  cudaMemset(plane, 0, bytes);
  printf("plane [%d][%d][%d] fp32 = %.1f MB\n", T, C, Rp, bytes / 1e6);
  for (int g = 296; g <= 4736; g *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaMemsetAsync(flush, it, flush_bytes);
      cudaEventRecord(e0);
      linread<<<g, 256>>>(reinterpret_cast<const float4*>(plane), n / 4, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;
    }
    printf("linear read, %4d CTAs x 4 float4 in flight per thread: %7.1f us  %6.2f TB/s\n", g, best * 1e3, bytes / best / 1e9);
  }
  const int splits[] = {1, 2, 4, 8, 16};
  for (int si = 0; si < 5; ++si) {
    const int s = splits[si];
    if (T / s < 8) continue;
    run<1, 2>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<1, 4>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<1, 8>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<1, 4>(plane, T, C, Rp, 1, s, out, bytes, flush, flush_bytes);
    run<2, 2>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<2, 4>(plane, T, C, Rp, 2, s, out, bytes, flush, flush_bytes);
    run<4, 1>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<4, 2>(plane, T, C, Rp, 4, s, out, bytes, flush, flush_bytes);
    run<4, 2>(plane, T, C, Rp, 2, s, out, bytes, flush, flush_bytes);
    run<4, 2>(plane, T, C, Rp, 1, s, out, bytes, flush, flush_bytes);
    run<4, 4>(plane, T, C, Rp, 1, s, out, bytes, flush, flush_bytes);
    run<8, 1>(plane, T, C, Rp, 2, s, out, bytes, flush, flush_bytes);
    run<8, 2>(plane, T, C, Rp, 1, s, out, bytes, flush, flush_bytes);
  }
  return 0;
}

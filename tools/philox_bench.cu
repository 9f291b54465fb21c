// philox_bench.cu — microbenchmark: how fast can one SM draw Philox4x32-R keep-masks, as a function
// of warps per scheduler and independent chains per thread?  (Sizing input for proj_tc.cu's producers.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o philox_bench philox_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define MCMIL_PHILOX_ROUNDS 10
#include "../montecarlo-gated-mil_b200/csrc/philox.cuh"
using namespace mcmil;

template <int R>
__device__ __forceinline__ uint4 philox_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey& key) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t keep_mask2(uint32_t r, uint32_t thr2) {
  uint32_t a, m;
  asm("abs.f16x2 %0, %1;" : "=r"(a) : "r"(r));
  asm("set.geu.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(thr2));
  return m;
}

template <int CHAINS, int R, bool MASK>
__global__ void bench_kernel(const __grid_constant__ PhiloxKey key, int iters, uint32_t thr2, uint32_t* out) {
  uint32_t acc = 0;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint4 h = make_uint4(tid * 3, tid * 5, tid * 7, tid * 11);
  for (int it = 0; it < iters; ++it) {
    uint4 r[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) r[c] = philox_r<R>((uint32_t)(it * CHAINS + c), tid, 17u, 3u, key);
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (MASK) {
        acc ^= (h.x & keep_mask2(r[c].x, thr2)) ^ (h.y & keep_mask2(r[c].y, thr2)) ^
               (h.z & keep_mask2(r[c].z, thr2)) ^ (h.w & keep_mask2(r[c].w, thr2));
      } else {
        acc ^= r[c].x ^ r[c].y ^ r[c].z ^ r[c].w;
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;   // keep the work alive
}

template <int CHAINS, int R, bool MASK>
void run(int warps, int sms, double mhz) {
  PhiloxKey key = philox_key(42);
  uint32_t* out; cudaMalloc(&out, 4);
  const int iters = 4096 / CHAINS;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_kernel<CHAINS, R, MASK><<<sms, warps * 32>>>(key, iters, 0x0CCD0CCDu, out);
  cudaEventRecord(e0);
  bench_kernel<CHAINS, R, MASK><<<sms, warps * 32>>>(key, iters, 0x0CCD0CCDu, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double calls_per_sm = (double)warps * iters * CHAINS;      // warp-level Philox calls per SM
  const double cycles = ms * 1e-3 * mhz * 1e6;
  printf("R=%2d mask=%d chains=%d warps/SM=%2d : %.1f cycles per warp-call per SMSP  (%.3f ms)\n", R, (int)MASK, CHAINS, warps,
         cycles / (calls_per_sm / 4.0), ms);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0;
  printf("%s, %d SMs, nominal %0.f MHz (cycles are computed at the nominal clock)\n", p.name, p.multiProcessorCount, mhz);
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8, 12, 16}) {
    run<1, 10, false>(w, sms, mhz); run<2, 10, false>(w, sms, mhz); run<4, 10, false>(w, sms, mhz); run<8, 10, false>(w, sms, mhz);
  }
  for (int w : {4, 8, 12, 16}) { run<4, 10, true>(w, sms, mhz); run<8, 10, true>(w, sms, mhz); }
  for (int w : {4, 8, 12, 16}) { run<4, 7, false>(w, sms, mhz); run<4, 7, true>(w, sms, mhz); }
  return 0;
}

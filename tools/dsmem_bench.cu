// dsmem_bench.cu — feasibility probe for sharing masked A slices between two CTA pairs of a 4-CTA
// cluster: (1) how many 4-CTA clusters with ~208 KB smem per CTA are co-resident on a B200,
// (2) bandwidth of cp.async.bulk shared::cta -> shared::cluster (TMA) between cluster peers,
// (3) bandwidth of st.shared::cluster.v4 (LSU path).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int SLICE = 8192;
constexpr int SMEM = 208 * 1024;

template <int MODE>   // 0: TMA bulk copy smem -> peer smem, 1: st.shared::cluster.v4 by 256 threads
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1) kern(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t rank = ctarank(), peer = rank ^ 2;
  const uint32_t bar = sb + SMEM - 64;          // mbarrier receiving the peer's copies
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * rank;
  __syncthreads();
  csync();
  const long long t0 = clock64();
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        // expect the 8 slices the peer sends me this round
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(8 * SLICE) : "memory");
        for (int s = 0; s < 8; ++s) {
          const uint32_t src = sb + s * SLICE;                       // my slices 0..7
          const uint32_t dst = mapa(sb + 65536 + s * SLICE, peer);   // peer's landing area
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(dst), "r"(src), "r"(SLICE), "r"(mapa(bar, peer)) : "memory");
        }
        uint32_t ok = 0, spins = 0;
        while (!ok && ++spins < (1u << 22)) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        if (!ok) { cycles[2] = it + 1; break; }      // protocol stuck: report instead of hanging
        phase ^= 1;
      }
    }
  } else {
    if (threadIdx.x < 256) {
      for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < 8; ++s) {
          const uint32_t off = s * SLICE + threadIdx.x * 16;
          const uint4 v = *reinterpret_cast<const uint4*>(smem + off);
          const uint32_t dst = mapa(sb + 65536 + off, peer);
          asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          const uint4 v2 = *reinterpret_cast<const uint4*>(smem + off + 4096);
          asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(dst + 4096), "r"(v2.x), "r"(v2.y), "r"(v2.z), "r"(v2.w) : "memory");
        }
      }
    }
  }
  __syncthreads();
  csync();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = clock64() - t0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s\n", p.name);
  for (int cs : {2, 4}) {
    cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)cs, 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(512); cfg.gridDim = dim3(cs * 37); cfg.dynamicSmemBytes = SMEM;
    cudaFuncSetAttribute(kern<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    int n = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern<0>, &cfg);
    printf("cluster size %d with %d KB smem/CTA: max active clusters = %d (%s)\n", cs, SMEM / 1024, n, cudaGetErrorString(e));
  }
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* cyc; cudaMalloc(&cyc, 32); cudaMemset(cyc, 0, 32);
  cudaFuncSetAttribute(kern<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  cudaFuncSetAttribute(kern<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int iters = 200;
  for (int grid : {4, 132}) {
    kern<1><<<grid, 512, SMEM>>>(iters, cyc); cudaError_t e1 = cudaDeviceSynchronize();
    printf("grid %d: LSU path done (%s)\n", grid, cudaGetErrorString(e1));
    kern<0><<<grid, 512, SMEM>>>(iters, cyc); cudaError_t e0 = cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    if (h[2]) printf("  TMA path stuck at iteration %lld\n", h[2]);
    const double bytes = (double)iters * 8 * SLICE;
    printf("grid %3d: TMA smem->peer smem  %.1f B/cycle per CTA (%s); st.shared::cluster.v4  %.1f B/cycle per CTA (%s)\n",
           grid, bytes / h[0], cudaGetErrorString(e0), bytes / h[1], cudaGetErrorString(e1));
  }
  return 0;
}

// pipe_bench.cu — issue-rate microbenchmark of the SASS instructions the producers use (sm_100a).
// Reports cycles per warp-instruction per SM sub-partition at saturation (16 warps/SM, 8 independent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEF_KERNEL(NAME, BODY)                                                                 \
  __global__ void NAME(int iters, uint32_t seed, uint32_t* out) {                              \
    uint32_t a[8], b = seed | 1u, c = seed * 3u + 7u;                                          \
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 17u + i * seed;                           \
    for (int it = 0; it < iters; ++it) {                                                       \
      _Pragma("unroll") for (int r = 0; r < 8; ++r) {                                          \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) { BODY }                                 \
      }                                                                                        \
    }                                                                                          \
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];                                     \
    if (s == 0x13572468u) out[0] = s;                                                          \
  }

DEF_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
DEF_KERNEL(k_iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0xBB99;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
DEF_KERNEL(k_imadhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_imadwide, { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b)); a[i] = (uint32_t)(w >> 32) ^ (uint32_t)w; })
DEF_KERNEL(k_hset2, asm volatile("set.geu.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
DEF_KERNEL(k_hmul2, asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
DEF_KERNEL(k_vsub2, asm volatile("sub.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_max, asm volatile("max.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_vmax2, asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
DEF_KERNEL(k_tanh, asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i]));)
// mixes
DEF_KERNEL(k_wide_lop3, { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b)); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((uint32_t)(w >> 32)), "r"((uint32_t)w), "r"(c)); })
DEF_KERNEL(k_wide_lop3_hset, { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b)); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((uint32_t)(w >> 32)), "r"((uint32_t)w), "r"(c)); if ((i & 3) == 0) asm volatile("set.geu.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); })

typedef void (*kern_t)(int, uint32_t, uint32_t*);
static void run(const char* name, kern_t k, int ops_per_inner, double mhz, int sms) {
  uint32_t* out; cudaMalloc(&out, 4);
  const int iters = 2000, warps = 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms, warps * 32>>>(iters, 12345u, out);
  cudaEventRecord(e0);
  k<<<sms, warps * 32>>>(iters, 12345u, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instrs_per_smsp = (double)iters * 64 * ops_per_inner * warps / 4.0;
  printf("%-18s %.2f cycles per warp-instruction per SMSP (%.3f ms)\n", name, ms * 1e-3 * mhz * 1e6 / warp_instrs_per_smsp, ms);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0; const int sms = p.multiProcessorCount;
  printf("%s nominal %.0f MHz\n", p.name, mhz);
  run("LOP3", k_lop3, 1, mhz, sms); run("IADD", k_iadd, 1, mhz, sms); run("PRMT", k_prmt, 1, mhz, sms);
  run("SHF", k_shf, 1, mhz, sms); run("IMAD.lo", k_imad, 1, mhz, sms); run("IMAD.HI", k_imadhi, 1, mhz, sms);
  run("IMAD.WIDE(+xor)", k_imadwide, 2, mhz, sms); run("HSET2", k_hset2, 1, mhz, sms); run("HFMA2", k_hfma2, 1, mhz, sms);
  run("HMUL2", k_hmul2, 1, mhz, sms); run("FFMA", k_ffma, 1, mhz, sms); run("ISUB", k_vsub2, 1, mhz, sms);
  run("UMAX", k_max, 1, mhz, sms); run("VMAX.U16x2", k_vmax2, 1, mhz, sms); run("MUFU.TANH", k_tanh, 1, mhz, sms);
  run("WIDE+LOP3", k_wide_lop3, 2, mhz, sms); run("WIDE+LOP3+HSET/4", k_wide_lop3_hset, 2, mhz, sms);
  return 0;
}

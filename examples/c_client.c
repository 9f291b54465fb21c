/* c_client.c — the drop-in boundary used from plain C: include/mcmil_b200.h + the CUDA runtime, no Python, no torch.
 *
 * What a non-Python host (or the reference-side binding of INTEGRATION.md) does for one call of the hot path
 * (model.py:256-328): upload the head's parameters in nn.Linear layout, describe the packed batch of bags, run all
 * T MC-dropout passes, read the per-sample logits and the Welford statistics back.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_client.c \
 *       -Lmontecarlo-gated-mil_b200/lib -lmcmil_b200 -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/montecarlo-gated-mil_b200/lib -o build/c_client
 *   build/c_client            -> prints the class probabilities of two bags and "c client ok"
 *
 * Checks (exit code 1 on failure): every output finite, the mean class probabilities of a bag sum to 1, the mean
 * attention of a head sums to 1 over the bag's patches, a second call with the same seed is bit-identical and the
 * tcgen05 path agrees with the fp32 CUDA-core path of the same library within the documented fp16-operand tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "mcmil_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)
#define CHECK_MCMIL(x) do { int r_ = (x); if (r_ != 0) { \
  fprintf(stderr, "%s:%d mcmil error %d: %s\n", __FILE__, __LINE__, r_, mcmil_last_error()); exit(2); } } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) {                       /* uniform in [-1, 1) */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) * (2.0f / 16777216.0f) - 1.0f;
}
static float* upload(const float* h, size_t n) {
  float* d = NULL;
  CHECK_CUDA(cudaMalloc((void**)&d, n * sizeof(float)));
  CHECK_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}
static float* filled(size_t n, float scale, int nonneg) {
  float* h = (float*)malloc(n * sizeof(float));
  for (size_t i = 0; i < n; ++i) { const float v = frand() * scale; h[i] = nonneg && v < 0.f ? 0.f : v; }
  return h;
}

int main(void) {
  enum { C = 2, T = 16, NB = 2 };
  const int32_t cu[NB + 1] = {0, 300, 300 + 77};         /* two bags: 300 and 77 patches */
  const int R = cu[NB];
  const float bound_in = 1.0f / sqrtf((float)MCMIL_L), bound_d = 1.0f / sqrtf((float)MCMIL_D);

  /* parameters, nn.Linear default-init ranges, shared attention (model.py:181-203) */
  float* hVw = filled((size_t)MCMIL_D * MCMIL_L, bound_in, 0); float* hVb = filled(MCMIL_D, bound_in, 0);
  float* hUw = filled((size_t)MCMIL_D * MCMIL_L, bound_in, 0); float* hUb = filled(MCMIL_D, bound_in, 0);
  float* hww = filled((size_t)C * MCMIL_D, bound_d, 0);        float* hwb = filled(C, bound_d, 0);
  float* hcw = filled((size_t)C * MCMIL_L, bound_in, 0);
  float* hH = filled((size_t)R * MCMIL_L, 1.0f, 1);            /* ReLU-like features */
  float *dVw = upload(hVw, (size_t)MCMIL_D * MCMIL_L), *dVb = upload(hVb, MCMIL_D);
  float *dUw = upload(hUw, (size_t)MCMIL_D * MCMIL_L), *dUb = upload(hUb, MCMIL_D);
  float *dww = upload(hww, (size_t)C * MCMIL_D), *dwb = upload(hwb, C), *dcw = upload(hcw, (size_t)C * MCMIL_L);
  float* dH = upload(hH, (size_t)R * MCMIL_L);

  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  mcmil_weights_t* w = NULL;
  mcmil_plan_t* plan = NULL;
  CHECK_MCMIL(mcmil_weights_create(&w, C, 1, dVw, dVb, dUw, dUb, dww, dwb, dcw, st));
  CHECK_MCMIL(mcmil_plan_create(&plan, cu, NULL, NB, T, C, st));
  if (mcmil_plan_total_rows(plan) != R) { fprintf(stderr, "plan rows\n"); return 1; }
  const size_t ws_bytes = mcmil_plan_workspace_bytes(plan);
  void* ws = NULL;
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));

  const size_t nY = (size_t)NB * T * C, nP = (size_t)NB * C, nA = (size_t)C * R;
  float *dY, *dPm, *dPq, *dAm, *dAq;
  CHECK_CUDA(cudaMalloc((void**)&dY, nY * 4)); CHECK_CUDA(cudaMalloc((void**)&dPm, nP * 4));
  CHECK_CUDA(cudaMalloc((void**)&dPq, nP * 4)); CHECK_CUDA(cudaMalloc((void**)&dAm, nA * 4));
  CHECK_CUDA(cudaMalloc((void**)&dAq, nA * 4));
  const size_t sz[5] = {nY, nP, nP, nA, nA};
  float* src[5];
  float* out[3][5];                                     /* [run][Y, prob_mean, prob_m2, attn_mean, attn_m2] */
  src[0] = dY; src[1] = dPm; src[2] = dPq; src[3] = dAm; src[4] = dAq;
  for (int run = 0; run < 3; ++run) {                   /* 0, 1: tcgen05 twice (same seed); 2: fp32 CUDA-core path */
    const int impl = run < 2 ? MCMIL_IMPL_TCGEN05 : MCMIL_IMPL_SIMT_FP32;
    CHECK_MCMIL(mcmil_head_forward(w, plan, dH, 0, 0, 7ull, 10, 0.1f, 0.1f, NULL, NULL, impl, dY, NULL, dPm, dPq, dAm,
                                   dAq, ws, ws_bytes, st));
    CHECK_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; ++k) {
      out[run][k] = (float*)malloc(sz[k] * 4);
      CHECK_CUDA(cudaMemcpy(out[run][k], src[k], sz[k] * 4, cudaMemcpyDeviceToHost));
    }
  }
  for (int k = 0; k < 5; ++k)                           /* same seed, same path: bit-identical */
    if (memcmp(out[0][k], out[1][k], sz[k] * 4) != 0) { fprintf(stderr, "run-to-run difference in output %d\n", k); return 1; }
  /* out[0]: tcgen05, out[2]: fp32 CUDA cores */
  int bad = 0;
  for (int b = 0; b < NB; ++b) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += out[0][1][b * C + c];
    printf("bag %d (%d patches): mean class probabilities", b, cu[b + 1] - cu[b]);
    for (int c = 0; c < C; ++c) printf(" %.5f (var %.2e)", out[0][1][b * C + c], out[0][2][b * C + c] / (T - 1));
    printf("\n");
    if (!(fabsf(s - 1.0f) < 1e-5f)) { fprintf(stderr, "probabilities of bag %d sum to %g\n", b, s); bad = 1; }
    for (int c = 0; c < C; ++c) {
      double a = 0.0;
      for (int n = cu[b]; n < cu[b + 1]; ++n) a += out[0][3][(size_t)c * R + n];
      if (!(fabs(a - 1.0) < 1e-4)) { fprintf(stderr, "attention of bag %d head %d sums to %g\n", b, c, a); bad = 1; }
    }
  }
  const float tol_rel[5] = {4e-3f, 3e-3f, 5e-2f, 4e-3f, 5e-2f};   /* fp16 operands vs fp32 (tests/test_gpu_parity.py REL) */
  for (int k = 0; k < 5; ++k) {
    float mx = 0.f, df = 0.f;
    for (size_t i = 0; i < sz[k]; ++i) {
      if (!isfinite(out[0][k][i])) { fprintf(stderr, "non-finite output %d[%zu]\n", k, i); return 1; }
      mx = fmaxf(mx, fabsf(out[2][k][i]));
      df = fmaxf(df, fabsf(out[0][k][i] - out[2][k][i]));
    }
    if (df > tol_rel[k] * mx + 1e-12f) { fprintf(stderr, "output %d: tcgen05 vs fp32 differ by %g (max %g)\n", k, df, mx); bad = 1; }
  }
  printf("kernel launches of the last call: %d\n", mcmil_last_launch_count());
  CHECK_MCMIL(mcmil_plan_destroy(plan));
  CHECK_MCMIL(mcmil_weights_destroy(w));
  if (bad) return 1;
  printf("c client ok\n");
  return 0;
}

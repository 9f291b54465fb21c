// pack.cu — layout kernels: weights -> UMMA shared-memory images (fp16, K-major, 128-byte
// swizzle) and the Philox keep-bit export used by the tests.
//
// Parameters packed here are the ones model.py:181-203 defines.
#include "internal.h"

namespace mcmil {

// byte offset of element (row r, k in [0,64)) inside one 128-byte-swizzled K-major slice
__host__ __device__ inline uint32_t sw128_off(int r, int k) {
  return (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

// ------------------------------------------------------------------ weights
// W image: [S][2 ranks][8 slices][136 rows][128 B] (row roles: common.cuh).  rank r, row j:
//   j <  32: tanh row    d = 32 r + j            32 <= j <  64: sigmoid row d = 32 r + (j - 32)
//   64 <= j < 72: classifier-score rows (rank 0 only; rows 0..3 = fp16 hi part, 4..7 = fp16(w - hi);
//                 shared: slot i <-> head i, separate set s: slot 0 <-> head s), rank 1: zeros
//   72 <= j < 104: tanh row d = 64 + 32 r + (j - 72)     104 <= j < 136: sigmoid row d = 64 + 32 r + (j - 104)
__global__ void pack_wmain_kernel(const float* __restrict__ V, const float* __restrict__ U,
                                  const float* __restrict__ cls, int S, int C, int shared,
                                  __half* __restrict__ img) {
  const int total = S * 2 * NSLICE * W_ROWS * KSLICE;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i % KSLICE;
    const int j = (i / KSLICE) % W_ROWS;
    const int q = (i / (KSLICE * W_ROWS)) % NSLICE;
    const int r = (i / (KSLICE * W_ROWS * NSLICE)) % 2;
    const int s = i / (KSLICE * W_ROWS * NSLICE * 2);
    const int k = q * KSLICE + kk;
    float v = 0.f;
    if (j < 64) {
      v = (j < 32 ? V : U)[((size_t)s * D + 32 * r + (j & 31)) * L + k];
    } else if (j < W_ROWS_A) {
      const int row = j - 64, slot = row & 3;
      const int head = shared ? slot : (slot == 0 ? s : -1);
      if (r == 0 && head >= 0 && head < C) {
        const float w = cls[(size_t)head * L + k];
        const float hi = __half2float(__float2half_rn(w));
        v = (row < 4) ? hi : (w - hi);
      }
    } else {
      const int jj = j - W_ROWS_A;
      v = (jj < 32 ? V : U)[((size_t)s * D + 64 + 32 * r + (jj & 31)) * L + k];
    }
    const size_t base = ((size_t)(s * 2 + r) * NSLICE + q) * SLICE_BYTES_W;
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(img) + base + sw128_off(j, kk)) = __float2half_rn(v);
  }
}

// fp32 transposed copy for the SIMT path: WT[s][k][j], j<128 -> V[s][j][k], else U[s][j-128][k]
__global__ void pack_wt_kernel(const float* __restrict__ V, const float* __restrict__ U, int S,
                               float* __restrict__ WT) {
  const int total = S * L * 256;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 256, k = (i / 256) % L, s = i / (256 * L);
    WT[i] = (j < 128 ? V : U)[((size_t)s * D + (j & 127)) * L + k];
  }
}

cudaError_t launch_pack_weights(Weights& w, const float* attV_w, const float* attV_b, const float* attU_w,
                                const float* attU_b, const float* attw_w, const float* attw_b,
                                const float* cls_w, cudaStream_t st) {
  const int S = w.S, C = w.C;
  pack_wmain_kernel<<<256, 256, 0, st>>>(attV_w, attU_w, cls_w, S, C, w.shared, reinterpret_cast<__half*>(w.d_wmain));
  pack_wt_kernel<<<256, 256, 0, st>>>(attV_w, attU_w, S, w.d_wt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  cudaMemcpyAsync(w.d_bv, attV_b, sizeof(float) * S * D, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(w.d_bu, attU_b, sizeof(float) * S * D, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(w.d_ww, attw_w, sizeof(float) * C * D, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(w.d_bw, attw_b, sizeof(float) * C, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(w.d_cls, cls_w, sizeof(float) * C * L, cudaMemcpyDeviceToDevice, st);
  // host copies of the epilogue constants (kernel-parameter constant bank)
  std::vector<float> bv(S * D), bu(S * D), ww(C * D), bw(C);
  cudaMemcpyAsync(bv.data(), attV_b, sizeof(float) * S * D, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(bu.data(), attU_b, sizeof(float) * S * D, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(ww.data(), attw_w, sizeof(float) * C * D, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(bw.data(), attw_b, sizeof(float) * C, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  w.epi.assign(S, EpiConst{});
  for (int s = 0; s < S; ++s) {
    EpiConst& ec = w.epi[s];
    for (int d = 0; d < D; d += 2)
      ec.vb[d / 2] = make_float4(bv[s * D + d], bv[s * D + d + 1], 0.5f * bu[s * D + d], 0.5f * bu[s * D + d + 1]);
    for (int c = 0; c < MAXC; ++c) {
      // shared: slot c <-> head c; separate: slot 0 <-> head s
      const int head = w.shared ? c : (c == 0 ? s : -1);
      const bool on = head >= 0 && head < C;
      for (int d = 0; d < D; d += 2)
        ec.hw[d / 2][c] = on ? make_float2(0.5f * ww[head * D + d], 0.5f * ww[head * D + d + 1]) : make_float2(0.f, 0.f);
      ec.bw[c] = on ? bw[head] : 0.f;
    }
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------ mask export (tests)
__global__ void export_feat_kernel(const int32_t* __restrict__ cu, const int32_t* __restrict__ row2bag,
                                   const int32_t* __restrict__ gbag, int R, int T, MaskSpec m,
                                   uint32_t* __restrict__ bits) {
  // one thread per (t, row, word of 32 features)
  const size_t total = (size_t)T * R * 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wd = (int)(i % 16);
    const int g = (int)((i / 16) % R);
    const int t = (int)(i / (16 * (size_t)R));
    const int b = row2bag[g];
    const uint32_t n = (uint32_t)(g - cu[b]);
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out |= feature_keep8(m.rounds, (uint32_t)(wd * 4 + j), n, (uint32_t)(m.t_offset + t), (uint32_t)(m.bag_offset + gbag[b]),
                           m.key, m.thr_f) << (8 * j);
    bits[i] = out;
  }
}
__global__ void export_attn_kernel(const int32_t* __restrict__ cu, const int32_t* __restrict__ row2bag,
                                   const int32_t* __restrict__ gbag, int R, int Rw, int T, int C, MaskSpec m,
                                   uint32_t* __restrict__ bits) {
  const size_t total = (size_t)T * C * Rw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wd = (int)(i % Rw);
    const int c = (int)((i / Rw) % C);
    const int t = (int)(i / ((size_t)Rw * C));
    uint32_t out = 0;
    for (int j = 0; j < 32; ++j) {
      const int g = wd * 32 + j;
      if (g >= R) break;
      const int b = row2bag[g];
      const uint4 r = attn_words_rt(m.rounds, (uint32_t)(c >> 2), (uint32_t)(g - cu[b]), (uint32_t)(m.t_offset + t),
                                 (uint32_t)(m.bag_offset + gbag[b]), m.key);
      out |= (attn_keep_from(r, c, m.thr_a) ? 1u : 0u) << j;
    }
    bits[i] = out;
  }
}

cudaError_t launch_export_masks(const Plan& p, const MaskSpec& m, uint32_t* feat_bits, uint32_t* attn_bits,
                                cudaStream_t st) {
  if (feat_bits) export_feat_kernel<<<1024, 256, 0, st>>>(p.d_cu, p.d_row2bag, p.d_gbag, p.R, p.T, m, feat_bits);
  if (attn_bits) export_attn_kernel<<<256, 256, 0, st>>>(p.d_cu, p.d_row2bag, p.d_gbag, p.R, p.Rw, p.T, p.C, m, attn_bits);
  return cudaGetLastError();
}

}  // namespace mcmil

// proj_simt.cu — fp32 CUDA-core implementation of the same projection as proj_tc.cu.
//
// Same inputs, same outputs (dropped+scaled logits and classifier scores, model.py:280-291 and
// the folded model.py:308-316), but every operand stays fp32 and tanh/sigmoid are the precise
// libdevice versions: it is the exact-precision cross-check of the tensor-core path on the GPU
// (MCMIL_IMPL_SIMT_FP32), at sizes where the CPU oracle is too slow.
//
// CTA = 64 patches x one MC sample x one (V,U) set; 256 threads, thread tile 8 rows x 8 columns
// (columns lane+32j: j<4 tanh branch, j>=4 the sigmoid branch of the same hidden unit).
#include "internal.h"

namespace mcmil {

constexpr int SM_ROWS = 64, SM_KSTEP = 32, SM_THREADS = 256, A_LD = SM_ROWS + 1;

struct SimtParams {
  const void* H;        // [R][512] fp32 or (h_f16) fp16
  int h_f16;
  const float* WT;      // this set: [512][256]
  const float* bv; const float* bu;   // this set: [128]
  const float* ww;      // [C][128]
  const float* bw;      // [C]
  const float* cls;     // [C][512]
  const TileDesc* tiles;
  float* logits; float* scores;
  const uint32_t* inj_feat; const uint32_t* inj_attn;
  int T, C, R, Rp, Rw, n_out, head0, t_offset, bag_offset, rounds;
  uint32_t thr_f, thr_a;
  float sf, sa;
  PhiloxKey key;
};

__global__ void __launch_bounds__(SM_THREADS)
proj_simt_kernel(const SimtParams P) {
  __shared__ float As[SM_KSTEP][A_LD];
  __shared__ float Bs[SM_KSTEP][256];
  const int tid = threadIdx.x, lane = tid & 31, ry = tid >> 5;
  const TileDesc td = P.tiles[blockIdx.x >> 1];
  const int half = blockIdx.x & 1;
  const int t = blockIdx.y;
  if (half * SM_ROWS >= td.nrows) return;
  const uint32_t tg = (uint32_t)(P.t_offset + t), bag = (uint32_t)(P.bag_offset + td.gbag);

  float acc[8][8];
  float sc[MAXC][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) sc[c][i] = 0.f;
  }
  const int lrow = tid >> 2, lchunk = tid & 3;          // loader mapping: 64 rows x 4 chunks of 8
  const int trow_l = half * SM_ROWS + lrow;

  for (int k0 = 0; k0 < L; k0 += SM_KSTEP) {
    // masked, scaled feature slice (model.py:281): As[kk][row]
    {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (trow_l < td.nrows) {
        const size_t e0 = (size_t)(td.row0 + trow_l) * L + k0 + lchunk * 8;
        if (P.h_f16) {
          const uint4 hv = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(P.H) + e0));
          const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h2[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
        } else {
          const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(P.H) + e0);
          const float4 a = __ldg(src), b = __ldg(src + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        const int q = k0 / 8 + lchunk;
        uint32_t bits;
        if (P.inj_feat == nullptr)
          bits = feature_keep8(P.rounds, (uint32_t)q, (uint32_t)(td.n0 + trow_l), tg, bag, P.key, P.thr_f);
        else
          bits = reinterpret_cast<const uint8_t*>(P.inj_feat)[((size_t)t * P.R + td.row0 + trow_l) * 64 + q];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = ((bits >> e) & 1u) ? v[e] * P.sf : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) As[lchunk * 8 + e][lrow] = v[e];
    }
    {
      const float4* src = reinterpret_cast<const float4*>(P.WT + (size_t)k0 * 256);
      float4* dst = reinterpret_cast<float4*>(&Bs[0][0]);
#pragma unroll
      for (int i = 0; i < (SM_KSTEP * 256 / 4) / SM_THREADS; ++i) dst[tid + i * SM_THREADS] = __ldg(src + tid + i * SM_THREADS);
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < SM_KSTEP; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ry * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // classifier score partials (model.py:308-316 folded): lane owns k = k0 + lane
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < P.n_out) {
        const float wck = __ldg(P.cls + (size_t)(P.head0 + c) * L + k0 + lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[c][i] = fmaf(As[lane][ry * 8 + i], wck, sc[c][i]);
      }
    }
    __syncthreads();
  }

  // gate + w-projection (model.py:285-290), reduced over the 32 lanes of the warp
  float part[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) part[c][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int d = lane + 32 * j;
    const float bvd = __ldg(P.bv + d), bud = __ldg(P.bu + d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float av = tanhf(acc[i][j] + bvd);
      const float au = 1.0f / (1.0f + expf(-(acc[i][j + 4] + bud)));
      const float g = av * au;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < P.n_out) part[c][i] = fmaf(g, __ldg(P.ww + (size_t)(P.head0 + c) * D + d), part[c][i]);
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        part[c][i] += __shfl_xor_sync(0xffffffffu, part[c][i], o);
        sc[c][i] += __shfl_xor_sync(0xffffffffu, sc[c][i], o);
      }
  if (lane < 8) {
    const int trow = half * SM_ROWS + ry * 8 + lane;
    if (trow < td.nrows) {
      const int g = td.row0 + trow;
      uint4 rnd = make_uint4(0, 0, 0, 0);
      if (P.inj_attn == nullptr) rnd = attn_words_rt(P.rounds, 0u, (uint32_t)(td.n0 + trow), tg, bag, P.key);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < P.n_out) {
          const int head = P.head0 + c;
          float pv = 0.f, sv = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { if (i == lane) { pv = part[c][i]; sv = sc[c][i]; } }
          float logit = pv + __ldg(P.bw + head);
          bool keep;
          if (P.inj_attn == nullptr) keep = attn_keep_from(rnd, head, P.thr_a);
          else keep = (P.inj_attn[((size_t)t * P.C + head) * P.Rw + (g >> 5)] >> (g & 31)) & 1u;
          logit = keep ? logit * P.sa : 0.f;
          const size_t o = ((size_t)t * P.C + head) * P.Rp + td.pcol0 + trow;
          P.logits[o] = logit;
          P.scores[o] = sv;
        }
      }
    }
  }
}

cudaError_t launch_proj_simt(const Weights& w, const Plan& p, const MaskSpec& m, const void* H, int h_f16,
                             float* logits, float* scores, cudaStream_t st, int* launches) {
  for (int s = 0; s < w.S; ++s) {
    SimtParams P;
    P.H = H; P.h_f16 = h_f16;
    P.WT = w.d_wt + (size_t)s * L * 256;
    P.bv = w.d_bv + s * D; P.bu = w.d_bu + s * D;
    P.ww = w.d_ww; P.bw = w.d_bw; P.cls = w.d_cls;
    P.tiles = p.d_tiles;
    P.logits = logits; P.scores = scores;
    P.inj_feat = m.inj_feat; P.inj_attn = m.inj_attn;
    P.T = p.T; P.C = p.C; P.R = p.R; P.Rp = p.Rp; P.Rw = p.Rw;
    P.n_out = w.shared ? w.C : 1;
    P.head0 = w.shared ? 0 : s;
    P.t_offset = m.t_offset; P.bag_offset = m.bag_offset;
    P.thr_f = m.thr_f; P.thr_a = m.thr_a; P.sf = m.sf; P.sa = m.sa;
    P.key = m.key;
    P.rounds = m.rounds;
    // blockIdx.y is limited to 65535: chunk T if ever needed
    for (int t0 = 0; t0 < p.T; t0 += 32768) {
      SimtParams Q = P;
      const int tn = (p.T - t0 < 32768) ? p.T - t0 : 32768;
      Q.t_offset = m.t_offset + t0;
      Q.logits = logits + (size_t)t0 * p.C * p.Rp;
      Q.scores = scores + (size_t)t0 * p.C * p.Rp;
      if (m.inj_feat) Q.inj_feat = m.inj_feat + (size_t)t0 * p.R * 16;
      if (m.inj_attn) Q.inj_attn = m.inj_attn + (size_t)t0 * p.C * p.Rw;
      proj_simt_kernel<<<dim3(p.n_tiles * 2, tn), SM_THREADS, 0, st>>>(Q);
      if (launches) ++*launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace mcmil

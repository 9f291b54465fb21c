// common.cuh — shared constants and device-side tables of the MC-dropout GA-MIL head.
// Math being implemented: /root/reference/model.py:280-316.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/mcmil_b200.h"

namespace mcmil {

constexpr int L = MCMIL_L;            // feature width            (model.py:139)
constexpr int D = MCMIL_D;            // attention hidden width   (model.py:140)
constexpr int MAXC = MCMIL_MAX_CLASSES;
constexpr int TILE_ROWS = 128;        // patches per CTA-pair tile (64 per CTA)
constexpr int HALF_ROWS = 64;
constexpr int KSLICE = 64;            // fp16 elements per 128-byte swizzle row
constexpr int NSLICE = L / KSLICE;    // 8 K-slices
constexpr int SLICE_BYTES_A = HALF_ROWS * 128;   // 8 KB : 64 rows x 128 B
// W slice of one CTA: 136 rows x 128 B.  rows [0,72) feed MMA "a" (N=144 over the pair):
// 32 tanh rows, the matching 32 sigmoid rows, 8 classifier-score rows; rows [72,136) feed MMA "b"
// (N=128): 32 tanh + 32 sigmoid rows of the upper hidden units.  (A separate N=16 score MMA costs
// ~48 tensor-pipe cycles per K=16 step instead of the 4 the math needs: profiles/r1.)
constexpr int W_ROWS_A = 72, W_ROWS_B = 64, W_ROWS = W_ROWS_A + W_ROWS_B;
constexpr int SLICE_BYTES_W = W_ROWS * 128;      // 17 KB

// One pair tile = up to 128 consecutive patches of one bag.
struct TileDesc {
  int bag;      // local bag index
  int n0;       // first patch (row within the bag)
  int row0;     // first packed row (cu[bag] + n0)
  int nrows;    // valid rows in this tile, 1..128
  int gbag;     // global bag id keying the Philox masks (bag_ids[bag], or bag)
  int pcol0;    // column of the tile's first patch in the logit / score planes (bag start padded to 32 columns)
  int pad_[2];
};

// Epilogue constants of one (V,U) parameter set, passed in the kernel-parameter constant bank.  Laid out per PAIR
// of hidden units (d, d+1), d even: the packed fp32x2 FMAs of the epilogue take their constant operand as one aligned
// 64-bit uniform register pair, so one 16-byte constant load feeds two FFMA2 with no register shuffling.
struct alignas(16) EpiConst {
  float4 vb[D / 2];            // (bv[d], bv[d+1], 0.5*bu[d], 0.5*bu[d+1])       (sigmoid(x) = 0.5*tanh(0.5x)+0.5)
  float2 hw[D / 2][MAXC];      // 0.5 * attention_weights[c].weight[d], [d+1]
  float bw[MAXC];              // attention_weights[c].bias
};

__host__ __device__ inline int drop_threshold(float p) {
  if (p <= 0.f) return 0;
  if (p >= 1.f) return 32768;
  return (int)(p * 32768.0f + 0.5f);
}
__host__ __device__ inline float drop_scale(float p) { return p >= 1.f ? 0.f : 1.0f / (1.0f - p); }

}  // namespace mcmil

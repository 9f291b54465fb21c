// attnmap.cu — MC statistics of the reconstructed attention maps.
//
// Replaces ImagePatcher.reconstruct_attention_map (image_patcher.py:83-110) followed by the mean /
// unbiased std over the MC passes (infer.py:212-219).  The reference paints every patch's attention
// into a (T,C,1,H,W) pixel tensor (3.5 GB at 2294x1914, T=100), divides by the overlap count and by
// each pass's own maximum, then reduces over T.  The maps are piecewise constant on the cells cut out
// by the tile boundaries, so everything is done per CELL (a few thousand) straight from the head's
// A[t][c][row] in HBM; pixels are a gather of the cell statistics.
#include "internal.h"

namespace mcmil {

constexpr int AM_THREADS = 256;

// one CTA per (t, c): cell value = mean of A over the covering patches (0 if none), per-pass maximum
__global__ void __launch_bounds__(AM_THREADS)
attnmap_cells_kernel(const float* __restrict__ A, int R, int row0, const int32_t* __restrict__ cell_ptr,
                     const int32_t* __restrict__ cell_idx, int n_cells, float* __restrict__ cellv,
                     float* __restrict__ vmax) {
  __shared__ float red[AM_THREADS / 32];
  const float* a = A + (size_t)blockIdx.x * R + row0;            // blockIdx.x = t*C + c
  float* out = cellv + (size_t)blockIdx.x * n_cells;
  float m = 0.f;
  for (int cell = threadIdx.x; cell < n_cells; cell += AM_THREADS) {
    const int p0 = cell_ptr[cell], p1 = cell_ptr[cell + 1];
    float s = 0.f;
    for (int p = p0; p < p1; ++p) s += a[cell_idx[p]];
    const float v = p1 > p0 ? s / (float)(p1 - p0) : 0.f;        // image_patcher.py:101-105
    out[cell] = v;
    m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
#pragma unroll
    for (int w = 1; w < AM_THREADS / 32; ++w) mm = fmaxf(mm, red[w]);
    vmax[blockIdx.x] = mm;                                       // image_patcher.py:106
  }
}

// one thread per (c, cell): Welford over t of cellv / vmax (image_patcher.py:107-108, infer.py:216-219)
__global__ void __launch_bounds__(AM_THREADS)
attnmap_welford_kernel(const float* __restrict__ cellv, const float* __restrict__ vmax, int T, int C,
                       int n_cells, float* __restrict__ mean_out, float* __restrict__ m2_out) {
  const int cell = blockIdx.x * AM_THREADS + threadIdx.x;
  const int c = blockIdx.y;
  if (cell >= n_cells) return;
  float mean = 0.f, m2 = 0.f;
  for (int t = 0; t < T; ++t) {
    const float x = cellv[((size_t)t * C + c) * n_cells + cell] / vmax[t * C + c];
    const float d = x - mean;
    mean += d / (float)(t + 1);
    m2 = fmaf(d, x - mean, m2);
  }
  mean_out[(size_t)c * n_cells + cell] = mean;
  m2_out[(size_t)c * n_cells + cell] = m2;
}

cudaError_t launch_attnmap(const float* A, int T, int C, int R, int row0, const int32_t* cell_ptr,
                           const int32_t* cell_idx, int n_cells, float* cellv, float* vmax, float* mean,
                           float* m2, cudaStream_t st) {
  attnmap_cells_kernel<<<T * C, AM_THREADS, 0, st>>>(A, R, row0, cell_ptr, cell_idx, n_cells, cellv, vmax);
  attnmap_welford_kernel<<<dim3((n_cells + AM_THREADS - 1) / AM_THREADS, C), AM_THREADS, 0, st>>>(
      cellv, vmax, T, C, n_cells, mean, m2);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ tiling: non-empty fraction per tile
// image_patcher.py:43-59: px_non_zero[i] = mean(tile[0] > 0) * 100 over channel 0 of every tile
__global__ void __launch_bounds__(AM_THREADS)
tile_nonzero_kernel(const float* __restrict__ img, int W, const int32_t* __restrict__ tiles, int patch,
                    float* __restrict__ pct) {
  __shared__ int red[AM_THREADS / 32];
  const int y0 = tiles[blockIdx.x * 6 + 0], x0 = tiles[blockIdx.x * 6 + 1];
  int cnt = 0;
  for (int i = threadIdx.x; i < patch * patch; i += AM_THREADS) {
    const int dy = i / patch, dx = i - dy * patch;
    cnt += img[(size_t)(y0 + dy) * W + x0 + dx] > 0.f ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < AM_THREADS / 32; ++w) s += red[w];
    pct[blockIdx.x] = (float)s / (float)(patch * patch) * 100.0f;
  }
}

// gather the selected tiles into the bag tensor (n, C, patch, patch)  (image_patcher.py:52, 117-128)
__global__ void gather_tiles_kernel(const float* __restrict__ img, int Cimg, int Himg, int W,
                                    const int32_t* __restrict__ tiles, const int32_t* __restrict__ sel, int patch,
                                    float* __restrict__ bag) {
  const int n = blockIdx.x, ch = blockIdx.y;
  const int tid = sel[n];
  const int y0 = tiles[tid * 6 + 0], x0 = tiles[tid * 6 + 1];
  const float* src = img + (size_t)ch * Himg * W;
  float* dst = bag + ((size_t)n * Cimg + ch) * patch * patch;
  for (int i = threadIdx.x; i < patch * patch; i += blockDim.x) {
    const int dy = i / patch, dx = i - dy * patch;
    dst[i] = src[(size_t)(y0 + dy) * W + x0 + dx];
  }
}

cudaError_t launch_tile_nonzero(const float* img, int W, const int32_t* tiles, int n_tiles, int patch, float* pct,
                                cudaStream_t st) {
  tile_nonzero_kernel<<<n_tiles, AM_THREADS, 0, st>>>(img, W, tiles, patch, pct);
  return cudaGetLastError();
}
cudaError_t launch_gather_tiles(const float* img, int Cimg, int Himg, int W, const int32_t* tiles,
                                const int32_t* sel, int n_sel, int patch, float* bag, cudaStream_t st) {
  if (n_sel == 0) return cudaSuccess;
  gather_tiles_kernel<<<dim3(n_sel, Cimg), 256, 0, st>>>(img, Cimg, Himg, W, tiles, sel, patch, bag);
  return cudaGetLastError();
}

}  // namespace mcmil

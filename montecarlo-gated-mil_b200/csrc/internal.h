// internal.h — host-side objects behind the opaque C-ABI handles and the kernel launchers.
#pragma once
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "philox.cuh"

namespace mcmil {

struct Weights {
  int C = 0;            // num_classes
  int shared = 1;
  int S = 1;            // number of (V,U) parameter sets
  // tcgen05 image, per set: [2 ranks][8 slices][136 rows x 128 B]
  uint8_t* d_wmain = nullptr;
  // fp32 copies for the SIMT path: WT [S][512][256] (k-major; cols 0..127 V, 128..255 U)
  float* d_wt = nullptr;
  float* d_bv = nullptr;   // [S][128]
  float* d_bu = nullptr;   // [S][128]
  float* d_ww = nullptr;   // [C][128]
  float* d_bw = nullptr;   // [C]
  float* d_cls = nullptr;  // [C][512]
  std::vector<EpiConst> epi;  // per set (host copy, passed by value to the kernel)
};

struct Plan {
  int n_bags = 0, T = 0, C = 0;
  int R = 0;        // total packed rows
  int Rp = 0;       // columns of the logit / score planes: every bag starts at a multiple of 32 columns
                    // (128-byte aligned segments: vector loads / cp.async in the reductions)
  int Rw = 0;       // 32-bit words per row of the injected logit keep-masks: ceil(R / 32)
  int n_tiles = 0;  // 128-row pair tiles
  int max_n = 0;
  int wsplit = 1;   // sample groups of the column-statistics kernel
  int sm_limit = 0; // SMs the projection kernel may use (0 = all): mcmil_plan_set_sm_limit
  std::vector<int32_t> cu;
  std::vector<TileDesc> tiles;
  int32_t* d_cu = nullptr;
  TileDesc* d_tiles = nullptr;
  int32_t* d_row2bag = nullptr;  // [R]
  int32_t* d_gbag = nullptr;     // [n_bags] global bag ids
  int32_t* d_pcol = nullptr;     // [n_bags] first plane column of each bag
  int4* d_cblk = nullptr;        // [n_cblk] column blocks of the column kernel: (plane column, packed row, bag, patches)
  int n_cblk = 0;
  // workspace layout (byte offsets)
  size_t off_logit = 0, off_score = 0, off_rowstat = 0, off_wpart = 0, off_wcount = 0, ws_bytes = 0;
};

struct MaskSpec {
  PhiloxKey key;
  uint32_t thr_f, thr_a;
  float sf, sa;          // 1/(1-p)
  int t_offset, bag_offset;
  int rounds;            // Philox rounds: 10 (default) or 7
  const uint32_t* inj_feat;   // [T][R][16] or null
  const uint32_t* inj_attn;   // [T][C][Rp/32] or null
};

// Programmatic dependent launch: the kernel may start (prologue, barrier / TMEM set-up, W loads) while
// its stream predecessor drains; it must execute griddepcontrol.wait before touching anything the
// predecessor wrote or still reads.  Hides ~2-4 us of launch latency per kernel in single-bag calls.
struct PdlLaunch {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  // pdl = false: a plain stream-ordered launch.  Used for SM-limited plans (several single-bag calls side by side on
  // private streams): a projection kernel launched early camps on its SMs in griddepcontrol.wait while the reduction
  // kernels of its own stream look for room on a GPU whose other SMs are held by the other streams' projections —
  // measured 45-50 us per bag with the attribute against 25-31 us without (profiles/r2_experiments.md §4).
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl = true) {
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool off = getenv("MCMIL_NO_PDL") != nullptr;     // A/B switch
    cfg.attrs = attr; cfg.numAttrs = (off || !pdl) ? 0 : 1;
  }
};

// ---- launchers (each returns a cudaError_t and bumps *launches) ----
cudaError_t launch_pack_weights(Weights& w, const float* attV_w, const float* attV_b, const float* attU_w,
                                const float* attU_b, const float* attw_w, const float* attw_b,
                                const float* cls_w, cudaStream_t st);
cudaError_t launch_proj_tc(const Weights& w, const Plan& p, const MaskSpec& m, const void* H, int h_f16,
                           float* logits, float* scores, float* dbg, cudaStream_t st, int* launches);
cudaError_t launch_proj_simt(const Weights& w, const Plan& p, const MaskSpec& m, const void* H, int h_f16,
                             float* logits, float* scores, cudaStream_t st, int* launches);
cudaError_t launch_reduce(const Plan& p, const float* logits, const float* scores, uint8_t* workspace,
                          float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean,
                          float* attn_m2, cudaStream_t st, int* launches);
int welford_split(int n_cblk, int C, int T);
int col_tiles_per_cta();
cudaError_t launch_export_masks(const Plan& p, const MaskSpec& m, uint32_t* feat_bits, uint32_t* attn_bits,
                                cudaStream_t st);
cudaError_t launch_attnmap(const float* A, int T, int C, int R, int row0, const int32_t* cell_ptr,
                           const int32_t* cell_idx, int n_cells, float* cellv, float* vmax, float* mean,
                           float* m2, cudaStream_t st);
cudaError_t launch_tile_nonzero(const float* img, int W, const int32_t* tiles, int n_tiles, int patch, float* pct,
                                cudaStream_t st);
cudaError_t launch_gather_tiles(const float* img, int Cimg, int Himg, int W, const int32_t* tiles,
                                const int32_t* sel, int n_sel, int patch, float* bag, cudaStream_t st);
cudaError_t launch_welford_pack(const float* mean, const float* m2, double count, int n, double* packed,
                                cudaStream_t st);
cudaError_t launch_aux_pairwise(const Plan& p, const float* A, int pos, int neg, int is_positive, float margin,
                                float scale, float eps, float* loss, cudaStream_t st);
cudaError_t launch_welford_unpack(const double* packed, int n, float* mean, float* m2, cudaStream_t st);

}  // namespace mcmil

// api.cu — the extern "C" boundary declared in include/mcmil_b200.h.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include <string>
#include <cstdlib>
#include <mutex>
#include "internal.h"

using namespace mcmil;

struct mcmil_weights : Weights {};
struct mcmil_plan : Plan {};

namespace {
thread_local std::string g_err;
thread_local int g_launches = 0;

// optional CUDA-event bracketing of the projection kernel(s), for bench.py's roofline figure.  One process-wide
// recorder guarded by a mutex (calls from several threads are serialised while it is on; off by default).
struct Profile {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  size_t used = 0;       // events recorded so far (pairs * 2)
  int kernels = 0;       // projection launches covered
} g_prof;
std::mutex g_prof_mu;

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}
template <class T>
cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)); }
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

MaskSpec make_mask_spec(int t_offset, int bag_offset, uint64_t seed, int rounds, float p_f, float p_a,
                        const uint32_t* inj_f, const uint32_t* inj_a) {
  MaskSpec m;
  m.key = philox_key(seed);
  m.thr_f = (uint32_t)drop_threshold(p_f);
  m.thr_a = (uint32_t)drop_threshold(p_a);
  m.sf = drop_scale(p_f);
  m.sa = drop_scale(p_a);
  m.t_offset = t_offset;
  m.bag_offset = bag_offset;
  m.rounds = rounds;
  m.inj_feat = inj_f;
  m.inj_attn = inj_a;
  return m;
}
}  // namespace

extern "C" {

const char* mcmil_last_error(void) { return g_err.c_str(); }
int mcmil_version(void) { return 100; }
int mcmil_last_launch_count(void) { return g_launches; }

int mcmil_weights_create(mcmil_weights_t** out, int num_classes, int shared_attention,
                         const float* attV_w, const float* attV_b, const float* attU_w, const float* attU_b,
                         const float* attw_w, const float* attw_b, const float* cls_w, void* stream) {
  if (!out || !attV_w || !attV_b || !attU_w || !attU_b || !attw_w || !attw_b || !cls_w)
    return fail(MCMIL_E_BADARG, "mcmil_weights_create: null pointer");
  if (num_classes < 1 || num_classes > MCMIL_MAX_CLASSES)
    return fail(MCMIL_E_UNSUPPORTED, "mcmil_weights_create: num_classes must be in [1,4]");
  mcmil_weights* w = new (std::nothrow) mcmil_weights();
  if (!w) return fail(MCMIL_E_NOMEM, "mcmil_weights_create: out of host memory");
  w->C = num_classes;
  w->shared = shared_attention ? 1 : 0;
  w->S = w->shared ? 1 : num_classes;
  cudaError_t e = cudaSuccess;
  const int S = w->S, C = w->C;
  if (e == cudaSuccess) e = dmalloc(&w->d_wmain, (size_t)S * 2 * NSLICE * SLICE_BYTES_W);
  if (e == cudaSuccess) e = dmalloc(&w->d_wt, (size_t)S * L * 256);
  if (e == cudaSuccess) e = dmalloc(&w->d_bv, (size_t)S * D);
  if (e == cudaSuccess) e = dmalloc(&w->d_bu, (size_t)S * D);
  if (e == cudaSuccess) e = dmalloc(&w->d_ww, (size_t)C * D);
  if (e == cudaSuccess) e = dmalloc(&w->d_bw, (size_t)C);
  if (e == cudaSuccess) e = dmalloc(&w->d_cls, (size_t)C * L);
  if (e == cudaSuccess)
    e = launch_pack_weights(*w, attV_w, attV_b, attU_w, attU_b, attw_w, attw_b, cls_w, (cudaStream_t)stream);
  if (e != cudaSuccess) { mcmil_weights_destroy(w); return cuda_fail(e, "mcmil_weights_create"); }
  *out = w;
  return 0;
}

int mcmil_weights_destroy(mcmil_weights_t* w) {
  if (!w) return 0;
  cudaFree(w->d_wmain); cudaFree(w->d_wt); cudaFree(w->d_bv); cudaFree(w->d_bu);
  cudaFree(w->d_ww); cudaFree(w->d_bw); cudaFree(w->d_cls);
  delete w;
  return 0;
}

int mcmil_plan_create(mcmil_plan_t** out, const int32_t* cu, const int32_t* bag_ids, int n_bags, int T,
                      int num_classes, void* stream) {
  if (!out || !cu) return fail(MCMIL_E_BADARG, "mcmil_plan_create: null pointer");
  if (n_bags < 1) return fail(MCMIL_E_BADARG, "mcmil_plan_create: need at least one bag");
  if (T < 1) return fail(MCMIL_E_BADARG, "mcmil_plan_create: T must be >= 1");
  if (num_classes < 1 || num_classes > MCMIL_MAX_CLASSES)
    return fail(MCMIL_E_UNSUPPORTED, "mcmil_plan_create: num_classes must be in [1,4]");
  if (cu[0] != 0) return fail(MCMIL_E_BADARG, "mcmil_plan_create: cu_seqlens[0] must be 0");
  for (int b = 0; b < n_bags; ++b)
    if (cu[b + 1] <= cu[b]) return fail(MCMIL_E_BADARG, "mcmil_plan_create: every bag needs at least one patch");
  mcmil_plan* p = new (std::nothrow) mcmil_plan();
  if (!p) return fail(MCMIL_E_NOMEM, "mcmil_plan_create: out of host memory");
  p->n_bags = n_bags; p->T = T; p->C = num_classes;
  p->cu.assign(cu, cu + n_bags + 1);
  p->R = cu[n_bags];
  p->Rw = (p->R + 31) / 32;
  std::vector<int32_t> row2bag((size_t)p->R), pcol((size_t)n_bags);
  std::vector<int4> cblk;
  const int tpc = col_tiles_per_cta();
  long long cols = 0;                        // every bag starts at a multiple of 32 plane columns (128 bytes)
  for (int b = 0; b < n_bags; ++b) {
    const int n = cu[b + 1] - cu[b];
    if (n > p->max_n) p->max_n = n;
    pcol[(size_t)b] = (int32_t)cols;
    for (int n0 = 0; n0 < n; n0 += TILE_ROWS) {
      TileDesc td{}; td.bag = b; td.n0 = n0; td.row0 = cu[b] + n0; td.gbag = bag_ids ? bag_ids[b] : b;
      td.nrows = (n - n0 < TILE_ROWS) ? n - n0 : TILE_ROWS;
      td.pcol0 = (int)cols + n0;
      p->tiles.push_back(td);
    }
    for (int r = cu[b]; r < cu[b + 1]; ++r) row2bag[(size_t)r] = b;
    for (int n0 = 0; n0 < n; n0 += tpc * TILE_ROWS)          // column blocks: up to tpc consecutive tiles of the bag
      cblk.push_back(make_int4((int)cols + n0, cu[b] + n0, b, n - n0 < tpc * TILE_ROWS ? n - n0 : tpc * TILE_ROWS));
    cols += (long long)align_up((size_t)n, 32);
  }
  if (cols > 0x7fffffffLL || (long long)T * num_classes * cols > (1LL << 40)) {
    delete p;
    return fail(MCMIL_E_UNSUPPORTED, "mcmil_plan_create: batch too large for one call");
  }
  if (cols > 0x7fffffffLL) { delete p; return fail(MCMIL_E_UNSUPPORTED, "mcmil_plan_create: batch too large for one call"); }
  p->Rp = (int)cols;
  p->n_tiles = (int)p->tiles.size();
  p->n_cblk = (int)cblk.size();
  p->wsplit = welford_split(p->n_cblk, p->C, T);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = dmalloc(&p->d_cu, (size_t)n_bags + 1);
  if (e == cudaSuccess) e = dmalloc(&p->d_tiles, (size_t)p->n_tiles);
  if (e == cudaSuccess) e = dmalloc(&p->d_row2bag, (size_t)p->R);
  if (e == cudaSuccess) e = dmalloc(&p->d_gbag, (size_t)n_bags);
  if (e == cudaSuccess) e = dmalloc(&p->d_pcol, (size_t)n_bags);
  if (e == cudaSuccess) e = dmalloc(&p->d_cblk, cblk.size());
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_cblk, cblk.data(), sizeof(int4) * cblk.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_pcol, pcol.data(), sizeof(int32_t) * n_bags, cudaMemcpyHostToDevice, st);
  std::vector<int32_t> gbag((size_t)n_bags);
  for (int b = 0; b < n_bags; ++b) gbag[(size_t)b] = bag_ids ? bag_ids[b] : b;
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_gbag, gbag.data(), sizeof(int32_t) * n_bags, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_cu, p->cu.data(), sizeof(int32_t) * (n_bags + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_tiles, p->tiles.data(), sizeof(TileDesc) * p->n_tiles, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_row2bag, row2bag.data(), sizeof(int32_t) * p->R, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // host staging buffers go out of scope
  if (e != cudaSuccess) { mcmil_plan_destroy(p); return cuda_fail(e, "mcmil_plan_create"); }
  size_t off = 0;
  p->off_logit = off;   off = align_up(off + (size_t)T * p->C * p->Rp * sizeof(float), 1024);
  p->off_score = off;   off = align_up(off + (size_t)T * p->C * p->Rp * sizeof(float), 1024);
  p->off_rowstat = off; off = align_up(off + (size_t)T * p->C * n_bags * sizeof(float2), 1024);
  p->off_wpart = off;   off = align_up(off + (p->wsplit > 1 ? (size_t)p->wsplit * p->C * p->Rp * sizeof(float2) : 0), 1024);
  p->off_wcount = off;  off = align_up(off + (size_t)p->n_cblk * p->C * sizeof(int), 1024);
  p->ws_bytes = off;
  *out = p;
  return 0;
}

int mcmil_plan_destroy(mcmil_plan_t* p) {
  if (!p) return 0;
  cudaFree(p->d_cu); cudaFree(p->d_tiles); cudaFree(p->d_row2bag); cudaFree(p->d_gbag); cudaFree(p->d_pcol); cudaFree(p->d_cblk);
  delete p;
  return 0;
}
size_t mcmil_plan_workspace_bytes(const mcmil_plan_t* p) { return p ? p->ws_bytes : 0; }
int mcmil_plan_total_rows(const mcmil_plan_t* p) { return p ? p->R : 0; }
int mcmil_plan_plane_cols(const mcmil_plan_t* p) { return p ? p->Rp : 0; }
int mcmil_plan_set_sm_limit(mcmil_plan_t* p, int sms) {
  if (!p || sms < 0) return fail(MCMIL_E_BADARG, "mcmil_plan_set_sm_limit: null plan or negative limit");
  p->sm_limit = sms;
  return 0;
}
int mcmil_plan_bag_plane_col(const mcmil_plan_t* p, int bag) {
  if (!p || bag < 0 || bag >= p->n_bags) return -1;
  int cols = 0;
  for (int b = 0; b < bag; ++b) cols += (int)align_up((size_t)(p->cu[(size_t)b + 1] - p->cu[(size_t)b]), 32);
  return cols;
}

static int head_forward_impl(const mcmil_weights_t* w, const mcmil_plan_t* plan, const void* H, int h_f16,
                             int t_offset, int bag_offset, uint64_t seed, int philox_rounds, float p_f, float p_a,
                             const uint32_t* inj_feat, const uint32_t* inj_attn, int impl,
                             float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean, float* attn_m2,
                             void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  if (!w || !plan || !H || !Y || !workspace) return fail(MCMIL_E_BADARG, "mcmil_head_forward: null pointer");
  if (w->C != plan->C) return fail(MCMIL_E_BADARG, "mcmil_head_forward: weights and plan disagree on num_classes");
  if ((inj_feat == nullptr) != (inj_attn == nullptr))
    return fail(MCMIL_E_BADARG, "mcmil_head_forward: inject both masks or neither");
  if (!(p_f >= 0.f && p_f <= 1.f) || !(p_a >= 0.f && p_a <= 1.f))
    return fail(MCMIL_E_BADARG, "mcmil_head_forward: dropout probability outside [0,1]");
  if (philox_rounds != 7 && philox_rounds != 10)
    return fail(MCMIL_E_BADARG, "mcmil_head_forward: philox_rounds must be 10 (default) or 7");
  if (workspace_bytes < plan->ws_bytes) return fail(MCMIL_E_WORKSPACE, "mcmil_head_forward: workspace too small");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023u) != 0)
    return fail(MCMIL_E_WORKSPACE, "mcmil_head_forward: workspace must be 1024-byte aligned");
  if ((reinterpret_cast<uintptr_t>(H) & 15u) != 0) return fail(MCMIL_E_BADARG, "mcmil_head_forward: H must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* logits = reinterpret_cast<float*>(ws + plan->off_logit);
  float* scores = reinterpret_cast<float*>(ws + plan->off_score);
  const MaskSpec m = make_mask_spec(t_offset, bag_offset, seed, philox_rounds, p_f, p_a, inj_feat, inj_attn);
  cudaError_t e;
  if (impl == MCMIL_IMPL_TCGEN05) {
    if (g_prof.on) {                       // measurement mode: bracket the projection launch(es) with events
      std::lock_guard<std::mutex> lock(g_prof_mu);
      const bool prof = g_prof.on && g_prof.used + 2 <= g_prof.ev.size();
      if (prof) cudaEventRecord(g_prof.ev[g_prof.used], st);
      const int before = g_launches;
      e = launch_proj_tc(*w, *plan, m, H, h_f16, logits, scores, nullptr, st, &g_launches);
      if (prof) { cudaEventRecord(g_prof.ev[g_prof.used + 1], st); g_prof.used += 2; g_prof.kernels += g_launches - before; }
    } else {
      e = launch_proj_tc(*w, *plan, m, H, h_f16, logits, scores, nullptr, st, &g_launches);
    }
    if (e != cudaSuccess) return cuda_fail(e, "proj_tc");
  } else if (impl == MCMIL_IMPL_SIMT_FP32) {
    e = launch_proj_simt(*w, *plan, m, H, h_f16, logits, scores, st, &g_launches);
    if (e != cudaSuccess) return cuda_fail(e, "proj_simt");
  } else {
    return fail(MCMIL_E_BADARG, "mcmil_head_forward: unknown impl");
  }
  e = launch_reduce(*plan, logits, scores, ws, Y, A, prob_mean, prob_m2, attn_mean, attn_m2, st, &g_launches);
  if (e != cudaSuccess) return cuda_fail(e, "reduce");
  return 0;
}

int mcmil_head_forward(const mcmil_weights_t* w, const mcmil_plan_t* plan, const float* H,
                       int t_offset, int bag_offset, uint64_t seed, int philox_rounds, float p_f, float p_a,
                       const uint32_t* inj_feat, const uint32_t* inj_attn, int impl,
                       float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean, float* attn_m2,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return head_forward_impl(w, plan, H, 0, t_offset, bag_offset, seed, philox_rounds, p_f, p_a, inj_feat, inj_attn, impl,
                           Y, A, prob_mean, prob_m2, attn_mean, attn_m2, workspace, workspace_bytes, stream);
}
int mcmil_head_forward_f16(const mcmil_weights_t* w, const mcmil_plan_t* plan, const uint16_t* H16,
                           int t_offset, int bag_offset, uint64_t seed, int philox_rounds, float p_f, float p_a,
                           const uint32_t* inj_feat, const uint32_t* inj_attn, int impl,
                           float* Y, float* A, float* prob_mean, float* prob_m2, float* attn_mean, float* attn_m2,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return head_forward_impl(w, plan, H16, 1, t_offset, bag_offset, seed, philox_rounds, p_f, p_a, inj_feat, inj_attn, impl,
                           Y, A, prob_mean, prob_m2, attn_mean, attn_m2, workspace, workspace_bytes, stream);
}

// Debug entry (not part of the reference-facing surface): runs pack + tcgen05 projection only and
// dumps each CTA's raw TMEM accumulators of its first (tile, t): dbg DEVICE fp32 [grid][128][136].
int mcmil_debug_proj_tc(const mcmil_weights_t* w, const mcmil_plan_t* plan, const float* H, int t_offset,
                        int bag_offset, uint64_t seed, float p_f, float p_a, const uint32_t* inj_feat,
                        const uint32_t* inj_attn, float* dbg, float* logits_out, float* scores_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!w || !plan || !H || !workspace) return fail(MCMIL_E_BADARG, "mcmil_debug_proj_tc: null pointer");
  if (workspace_bytes < plan->ws_bytes) return fail(MCMIL_E_WORKSPACE, "mcmil_debug_proj_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* logits = reinterpret_cast<float*>(ws + plan->off_logit);
  float* scores = reinterpret_cast<float*>(ws + plan->off_score);
  const MaskSpec m = make_mask_spec(t_offset, bag_offset, seed, 10, p_f, p_a, inj_feat, inj_attn);
  int launches = 0;
  cudaError_t e = launch_proj_tc(*w, *plan, m, H, 0, logits, scores, dbg, st, &launches);
  const size_t plane = (size_t)plan->T * plan->C * plan->Rp * sizeof(float);
  if (e == cudaSuccess && logits_out) e = cudaMemcpyAsync(logits_out, logits, plane, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && scores_out) e = cudaMemcpyAsync(scores_out, scores, plane, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "mcmil_debug_proj_tc");
  return 0;
}

int mcmil_profile_begin(int max_calls) {
  if (max_calls < 1) return fail(MCMIL_E_BADARG, "mcmil_profile_begin: max_calls must be >= 1");
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.assign((size_t)max_calls * 2, nullptr);
  for (auto& e : g_prof.ev) {
    cudaError_t err = cudaEventCreate(&e);
    if (err != cudaSuccess) return cuda_fail(err, "mcmil_profile_begin");
  }
  g_prof.used = 0; g_prof.kernels = 0; g_prof.on = true;
  return 0;
}
int mcmil_profile_end(double* total_ms, int* kernels) {
  if (!total_ms || !kernels) return fail(MCMIL_E_BADARG, "mcmil_profile_end: null pointer");
  std::lock_guard<std::mutex> lock(g_prof_mu);
  double total = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    cudaError_t err = cudaEventSynchronize(g_prof.ev[i + 1]);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]);
    if (err != cudaSuccess) return cuda_fail(err, "mcmil_profile_end");
    total += ms;
  }
  *total_ms = total; *kernels = g_prof.kernels;
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear(); g_prof.used = 0; g_prof.on = false;
  return 0;
}

int mcmil_attnmap_stats(const float* A, int T, int C, int R, int row0, const int32_t* cell_ptr,
                        const int32_t* cell_idx, int n_cells, float* cellv_ws, float* vmax_ws, float* cell_mean,
                        float* cell_m2, void* stream) {
  if (!A || !cell_ptr || !cell_idx || !cellv_ws || !vmax_ws || !cell_mean || !cell_m2)
    return fail(MCMIL_E_BADARG, "mcmil_attnmap_stats: null pointer");
  if (T < 1 || C < 1 || n_cells < 1 || row0 < 0 || R < 1) return fail(MCMIL_E_BADARG, "mcmil_attnmap_stats: bad size");
  cudaError_t e = launch_attnmap(A, T, C, R, row0, cell_ptr, cell_idx, n_cells, cellv_ws, vmax_ws, cell_mean,
                                 cell_m2, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_attnmap_stats");
}

int mcmil_tile_nonzero_pct(const float* image, int W, const int32_t* tiles, int n_tiles, int patch, float* pct,
                           void* stream) {
  if (!image || !tiles || !pct || n_tiles < 1 || patch < 1 || W < patch)
    return fail(MCMIL_E_BADARG, "mcmil_tile_nonzero_pct: bad argument");
  cudaError_t e = launch_tile_nonzero(image, W, tiles, n_tiles, patch, pct, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_tile_nonzero_pct");
}

int mcmil_gather_tiles(const float* image, int channels, int H, int W, const int32_t* tiles, const int32_t* selected,
                       int n_selected, int patch, float* bag, void* stream) {
  if (!image || !tiles || !selected || !bag || n_selected < 0 || channels < 1)
    return fail(MCMIL_E_BADARG, "mcmil_gather_tiles: bad argument");
  cudaError_t e = launch_gather_tiles(image, channels, H, W, tiles, selected, n_selected, patch, bag,
                                      (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_gather_tiles");
}

int mcmil_welford_pack(const float* mean, const float* m2, double count, int n, double* packed, void* stream) {
  if (!mean || !m2 || !packed || n < 0) return fail(MCMIL_E_BADARG, "mcmil_welford_pack: bad argument");
  cudaError_t e = launch_welford_pack(mean, m2, count, n, packed, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_welford_pack");
}
int mcmil_welford_unpack(const double* packed, int n, float* mean, float* m2, void* stream) {
  if (!mean || !m2 || !packed || n < 0) return fail(MCMIL_E_BADARG, "mcmil_welford_unpack: bad argument");
  cudaError_t e = launch_welford_unpack(packed, n, mean, m2, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_welford_unpack");
}

int mcmil_aux_pairwise_loss(const mcmil_plan_t* plan, const float* A, int T, int R, int pos_head, int neg_head,
                            int is_positive, float margin, float scale, float eps, float* loss, void* stream) {
  if (!plan || !A || !loss) return fail(MCMIL_E_BADARG, "mcmil_aux_pairwise_loss: null argument");
  if (T != plan->T || R != plan->R)
    return fail(MCMIL_E_BADARG, "mcmil_aux_pairwise_loss: A is not the (T, C, R) attention of this plan");
  if (pos_head < 0 || pos_head >= plan->C || neg_head < 0 || neg_head >= plan->C)
    return fail(MCMIL_E_BADARG, "mcmil_aux_pairwise_loss: head index out of range");
  cudaError_t e = launch_aux_pairwise(*plan, A, pos_head, neg_head, is_positive, margin, scale, eps, loss, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_aux_pairwise_loss");
}

int mcmil_export_masks(const mcmil_plan_t* plan, int t_offset, int bag_offset, uint64_t seed, int philox_rounds,
                       float p_f, float p_a, uint32_t* feat_bits, uint32_t* attn_bits, void* stream) {
  if (!plan) return fail(MCMIL_E_BADARG, "mcmil_export_masks: null plan");
  if (philox_rounds != 7 && philox_rounds != 10) return fail(MCMIL_E_BADARG, "mcmil_export_masks: philox_rounds must be 10 or 7");
  const MaskSpec m = make_mask_spec(t_offset, bag_offset, seed, philox_rounds, p_f, p_a, nullptr, nullptr);
  cudaError_t e = launch_export_masks(*plan, m, feat_bits, attn_bits, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "mcmil_export_masks");
}

}  // extern "C"

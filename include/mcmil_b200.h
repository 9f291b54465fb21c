/*
 * mcmil_b200.h — C ABI of the B200-native MC-dropout gated-attention MIL head.
 *
 * The reference (xkuubix/MonteCarlo-Gated-MIL) is pure Python/PyTorch and has no FFI,
 * plugin or operator registry: its boundary for this path is the
 * torch.nn.Module method
 *     MultiHeadGatedAttentionMIL.mc_inference(input_tensor, N, device)   model.py:256-328
 * whose head part (model.py:280-316) is what this library replaces.  The entry points
 * below are what a binding of that method would call; the Python side
 * (montecarlo-gated-mil_b200/head.py) binds them with ctypes and keeps the reference's
 * module interface (same constructor, same state_dict keys, same return tuple).
 *
 * Conventions
 *   - plain C, no torch types; every pointer documented as HOST or DEVICE;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the host (graph-capturable) unless stated;
 *   - every function returns 0 on success, a negative MCMIL_E* code for bad arguments,
 *     or a positive cudaError_t; mcmil_last_error() returns a message for the calling
 *     thread's last failure;
 *   - L (feature width) = 512 and D (attention hidden width) = 128 as in the reference
 *     defaults (model.py:139-140); num_classes in [1, 4].
 */
#ifndef MCMIL_B200_H_
#define MCMIL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCMIL_L 512
#define MCMIL_D 128
#define MCMIL_MAX_CLASSES 4

#define MCMIL_E_BADARG   (-1)
#define MCMIL_E_NOMEM    (-2)
#define MCMIL_E_WORKSPACE (-3)
#define MCMIL_E_UNSUPPORTED (-4)

/* projection implementation selector (mcmil_head_forward `impl`) */
#define MCMIL_IMPL_TCGEN05 0   /* fp16 operands, fp32 TMEM accumulators, sm_100a tensor cores */
#define MCMIL_IMPL_SIMT_FP32 1 /* fp32 CUDA-core path: exact-precision cross-check          */

typedef struct mcmil_weights mcmil_weights_t;
typedef struct mcmil_plan mcmil_plan_t;

const char* mcmil_last_error(void);
int mcmil_version(void);

/* ---- weights: replaces the parameter set of model.py:181-203 -------------------------
 * All pointers DEVICE fp32 in nn.Linear layout (row-major (out,in)), S = 1 if shared
 * else num_classes:
 *   attV_w [S][128][512], attV_b [S][128]     attention_V(.c).0.{weight,bias}  model.py:183,186
 *   attU_w [S][128][512], attU_b [S][128]     attention_U(.c).0.{weight,bias}  model.py:184,190
 *   attw_w [C][128],      attw_b [C]          attention_weights.c.{weight,bias} model.py:196
 *   cls_w  [C][512]                           classifiers.c.weight (no bias)    model.py:201
 * Repacks them privately (fp16 UMMA smem images, fp32 transposed copies). */
int mcmil_weights_create(mcmil_weights_t** out, int num_classes, int shared_attention,
                         const float* attV_w, const float* attV_b,
                         const float* attU_w, const float* attU_b,
                         const float* attw_w, const float* attw_b,
                         const float* cls_w, void* stream);
int mcmil_weights_destroy(mcmil_weights_t* w);

/* ---- plan: shapes of one call (the reference is bs==1, model.py:309; a plan may also
 * describe a packed variable-length batch of bags) ------------------------------------
 *   cu_seqlens_host  HOST int32 [n_bags+1], cu[0]=0, bag b owns packed rows [cu[b],cu[b+1])
 *   bag_ids_host     HOST int32 [n_bags] or NULL: global id of each bag (keys the RNG, so a
 *                    bag's masks do not depend on which rank/batch it lands in); NULL = 0..n_bags-1
 *   T                MC samples computed by THIS call (a shard of the job's samples)
 * The plan owns a small device copy of the bag/tile tables. */
int mcmil_plan_create(mcmil_plan_t** out, const int32_t* cu_seqlens_host, const int32_t* bag_ids_host,
                      int n_bags, int T, int num_classes, void* stream);
int mcmil_plan_destroy(mcmil_plan_t* p);
size_t mcmil_plan_workspace_bytes(const mcmil_plan_t* p);
int mcmil_plan_total_rows(const mcmil_plan_t* p);
/* Layout of the internal logit / score planes [T][C][plane_cols] (what mcmil_debug_proj_tc copies out): every bag
 * starts at a multiple of 32 columns; mcmil_plan_bag_plane_col returns the first column of a bag (-1: bad index). */
int mcmil_plan_plane_cols(const mcmil_plan_t* p);
/* Limits the projection kernel of calls with this plan to `sms` streaming multiprocessors (0 = all; rounded down to
 * CTA pairs, at least one).  Several single-bag calls issued on DIFFERENT streams then run side by side, each on its
 * share of the GPU, and the fixed per-kernel cost of one call (prologue, pipeline fill, tail: ~10 of ~33 us for one
 * bag of 1024 patches, T = 100) overlaps with the steady state of the others.  Results do not depend on the limit.
 * Calls with a limited plan are launched without programmatic dependent launch (see internal.h, PdlLaunch). */
int mcmil_plan_set_sm_limit(mcmil_plan_t* p, int sms);
int mcmil_plan_bag_plane_col(const mcmil_plan_t* p, int bag);

/* ---- the hot path: model.py:280-316 + the MC statistics of infer.py:195,212-219 -------
 *   H            DEVICE fp32 [R][512] packed patch features (R = cu[n_bags])
 *   t_offset     global index of this call's first MC sample (MC-sample sharding)
 *   bag_offset   added to every bag id — both offsets only key the RNG
 *   seed         Philox key; masks are a pure function of (seed, bag, t, n, l)
 *   philox_rounds  10 = Philox4x32-10 (what ATen / cuRAND use; default), 7 = Philox4x32-7 (the
 *                smallest Crush-resistant round count; ~25 % faster projection kernel)
 *   p_f, p_a     feature / logit dropout probabilities (model.py:141-142)
 *   inj_feat_keep_bits  DEVICE uint32 [T][R][16]   nullable; bit l%32 of word l/32: 1 = keep
 *   inj_attn_keep_bits  DEVICE uint32 [T][C][ceil(R/32)] nullable; bit r%32 of word r/32 (r = packed row)
 *                (both or neither; when given they replace the Philox masks — this is how the
 *                 reference's own masks are injected for bit-exact comparison)
 *   impl         MCMIL_IMPL_*
 * outputs (DEVICE fp32 unless noted; nullable ones are skipped when NULL)
 *   Y            [n_bags][T][C]   per-sample logits  (model.py:313-316; (T,1,C) for one bag)
 *   A            [T][C][R]        per-sample attention, nullable (model.py:305; (T,1,C,N))
 *   prob_mean, prob_m2   [n_bags][C]  Welford over the T samples of softmax_c(Y)
 *   attn_mean, attn_m2   [C][R]       Welford over the T samples of A   (count = T)
 *   workspace    DEVICE, >= mcmil_plan_workspace_bytes(plan), 1024-byte aligned
 */
int mcmil_head_forward(const mcmil_weights_t* w, const mcmil_plan_t* plan, const float* H,
                       int t_offset, int bag_offset, uint64_t seed, int philox_rounds, float p_f, float p_a,
                       const uint32_t* inj_feat_keep_bits, const uint32_t* inj_attn_keep_bits,
                       int impl, float* Y, float* A, float* prob_mean, float* prob_m2,
                       float* attn_mean, float* attn_m2, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Same call for features that already are IEEE half precision (e.g. an extractor run under autocast): H16 is DEVICE
 * fp16 [R][512].  The tensor-core path rounds fp32 features to fp16 anyway, so for H16 = fp16(H) both entry points
 * return bit-identical results; this one moves half the bytes (host-to-device copies included). */
int mcmil_head_forward_f16(const mcmil_weights_t* w, const mcmil_plan_t* plan, const uint16_t* H16,
                           int t_offset, int bag_offset, uint64_t seed, int philox_rounds, float p_f, float p_a,
                           const uint32_t* inj_feat_keep_bits, const uint32_t* inj_attn_keep_bits,
                           int impl, float* Y, float* A, float* prob_mean, float* prob_m2,
                           float* attn_mean, float* attn_m2, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---- Welford merge across MC-sample shards ----------------------------
 * pack:   out[0] = count, out[1+i] = count*mean[i], out[1+n+i] = m2[i] + count*mean[i]^2
 *         (DEVICE fp64 [1+2n]; additive, so ONE allreduce(sum) merges all shards)
 * unpack: inverse, writes merged mean / m2 (DEVICE fp32 [n]) and count to *count_host via
 *         a DEVICE double the caller reads (packed[0]). */
int mcmil_welford_pack(const float* mean, const float* m2, double count, int n, double* packed,
                       void* stream);
int mcmil_welford_unpack(const double* packed, int n, float* mean, float* m2, void* stream);

/* ---- mask export: the keep-bits the Philox path draws, for tests and for feeding the
 * reference module the very same masks --------------------------------------------------
 *   feat_bits DEVICE uint32 [T][R][16], attn_bits DEVICE uint32 [T][C][ceil(R/32)] */
int mcmil_export_masks(const mcmil_plan_t* plan, int t_offset, int bag_offset, uint64_t seed,
                       int philox_rounds, float p_f, float p_a, uint32_t* feat_bits, uint32_t* attn_bits,
                       void* stream);

/* ---- debug (tests only; not a reference-facing entry point): feature packing + tcgen05
 * projection only; dumps every CTA's raw TMEM accumulators of its first (tile, sample):
 *   dbg DEVICE fp32 [grid][128 lanes][136] (128 accumulator columns + 8 score columns), and the
 *   dropped logits / classifier scores planes [T][C][mcmil_plan_plane_cols(plan)]. */
int mcmil_debug_proj_tc(const mcmil_weights_t* w, const mcmil_plan_t* plan, const float* H, int t_offset,
                        int bag_offset, uint64_t seed, float p_f, float p_a,
                        const uint32_t* inj_feat_keep_bits, const uint32_t* inj_attn_keep_bits,
                        float* dbg, float* logits_out, float* scores_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- "next" rows of SURVEY.md §8f: the consumer and the producer either side of the head ----------
 * mcmil_attnmap_stats replaces ImagePatcher.reconstruct_attention_map (image_patcher.py:83-110) + the
 * mean / std over the MC passes (infer.py:212-219), at tile-boundary CELL resolution:
 *   A          DEVICE fp32 [T][C][R] (mcmil_head_forward's A), this bag's rows start at row0
 *   cell_ptr   DEVICE int32 [n_cells+1], cell_idx DEVICE int32 [nnz]: CSR of the bag positions
 *              (0..n-1) of the selected patches that cover each cell
 *   cellv_ws   DEVICE fp32 [T][C][n_cells], vmax_ws DEVICE fp32 [T][C]   (workspace)
 *   cell_mean / cell_m2  DEVICE fp32 [C][n_cells]: Welford over t of the per-pass max-normalised
 *              overlap-averaged attention of the cell (count = T; std = sqrt(m2/(T-1)))
 * mcmil_tile_nonzero_pct / mcmil_gather_tiles replace the loop of ImagePatcher.convert_img_to_bag
 * (image_patcher.py:43-59): per-tile percentage of channel-0 pixels > 0, and the gather of the selected
 * tiles (rows of `tiles` = (y, x, h, w, i, j) as image_patcher.py:36) into a (n, channels, patch, patch) bag. */
int mcmil_attnmap_stats(const float* A, int T, int C, int R, int row0, const int32_t* cell_ptr,
                        const int32_t* cell_idx, int n_cells, float* cellv_ws, float* vmax_ws, float* cell_mean,
                        float* cell_m2, void* stream);
int mcmil_tile_nonzero_pct(const float* image /*[channels][H][W], channel 0 is read*/, int W, const int32_t* tiles,
                           int n_tiles, int patch, float* pct, void* stream);
int mcmil_gather_tiles(const float* image, int channels, int H, int W, const int32_t* tiles, const int32_t* selected,
                       int n_selected, int patch, float* bag, void* stream);

/* ---- deterministic forward + auxiliary loss ----------------------------------------
 * The eval-mode forward of the head (model.py:211-253: no dropout, one pass) is mcmil_head_forward with
 * T = 1 and p_f = p_a = 0 (threshold 0 keeps every element; Y[.][0][:] / A[0] are the outputs).
 * mcmil_aux_pairwise_loss replaces AuxiliaryLoss.pairwise_distance_loss (model.py:405-426) as the model
 * applies it per MC pass (model.py:318-326) or once (model.py:243-248):
 *   A     DEVICE fp32 [T][C][R] (mcmil_head_forward's A); T and R must be the plan's (checked)
 *   loss  DEVICE fp32 [n_bags][T]:  scale * (is_positive ? max(margin - d, 0) : d),
 *         d = || A[t][pos_head][bag rows] - A[t][neg_head][bag rows] + eps ||_2   (F.pairwise_distance)
 * The reference uses pos_head = 1, neg_head = 0, margin = 1.0, scale = 0.5 (model.py:149-151), eps = 1e-6. */
int mcmil_aux_pairwise_loss(const mcmil_plan_t* plan, const float* A, int T, int R, int pos_head, int neg_head,
                            int is_positive, float margin, float scale, float eps, float* loss, void* stream);

/* ---- measurement hook (bench.py): brackets the tcgen05 projection launch(es) of the next
 * `max_calls` mcmil_head_forward calls with CUDA events on the launching stream;
 * mcmil_profile_end synchronises them and returns the summed device time and the number of
 * projection kernels covered.  One process-wide recorder (mutex-guarded); off by default. */
int mcmil_profile_begin(int max_calls);
int mcmil_profile_end(double* total_ms, int* kernels);

/* number of kernels the last mcmil_head_forward on this thread launched (bench "gpu_launches") */
int mcmil_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MCMIL_B200_H_ */

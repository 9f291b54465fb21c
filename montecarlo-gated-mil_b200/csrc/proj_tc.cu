// proj_tc.cu — the fused MC-dropout gated-attention projection on sm_100a tensor cores.
//
// Computes, for every MC sample t and patch n of a packed batch of bags (model.py:280-291,
// and the classifier contraction of model.py:308-316 folded in as extra output columns):
//     Hd[t,n,:]  = H[n,:] * keep_f[t,n,:]                      (1/(1-p_f) applied to the accumulators)
//     logit[t,c,n] = ( sum_d tanh(Hd Wv^T + bv)_d * sigmoid(Hd Wu^T + bu)_d * w_c[d] + b_c ) * keep_a / (1-p_a)
//     score[t,c,n] = Hd[t,n,:] . wc_c / (1-p_f)
//
// One persistent 2-CTA cluster per SM pair; the pair shares tcgen05.mma.cta_group::2 tiles of
// M=128 (64 patches per CTA) x N=144+128 (tanh | sigmoid | score columns, half of the W rows
// resident in each CTA's smem) x K=512.  Everything re-used across the T samples stays on chip:
//   smem : W (fp16, 136 KB/CTA, loaded once per kernel by TMA bulk copies) and an 8-slot ring of
//          masked fp16 A slices (slot = K-slice, 64 KB/CTA, a full sample deep);
//   regs : each producer thread keeps ITS 16 chunks of the fp16 feature tile in registers for the
//          whole work item (converted from the fp32 features once per item);
//   TMEM : three accumulator buffers (sample t accumulates into buffer t % 3 while the epilogue drains t-1, t-2).
// Warp roles (16 warps with 2 producer teams): 0-3 epilogue (TMEM -> tanh/sigmoid/gate/w-dot -> logits),
//   4-11 producers (Philox mask -> masked fp16 A slice, generic-proxy st.shared + proxy fence),
//   12-15 MMA issuers (sample tc is issued by warp 12 + tc % 4; 12 owns the TMEM allocation, 13 first
//   TMA-loads W).  One issuer per SM sub-partition: issuing a sample's 64 tcgen05.mma costs its
//   sub-partition ~25 % of a sample time, and with two issuers the two producer warps sharing their
//   sub-partitions were the laggards every K-slice barrier waited for (profiles/r1_experiments.md).
#include <cstdio>
#include <mutex>
#include "internal.h"
#include "ptx.cuh"

namespace mcmil {
using namespace ptx;

#ifndef MCMIL_TEAMS
#define MCMIL_TEAMS 2   // the mask / cache indexing of the Philox modes assumes 2 (static_asserts below)
#endif
constexpr int TEAMS = MCMIL_TEAMS;               // producer teams: team k fills the K-slices s = k (mod TEAMS)
constexpr int TEAM_WARPS = 4;                    // warps per team (16 patch rows each)
constexpr int TEAM_SLICES = NSLICE / TEAMS;      // slices per team and sample
#ifndef MCMIL_MMA_WARPS
#define MCMIL_MMA_WARPS 4   // MMA-issue warps, one per SM sub-partition of the leader CTA (see the role comment below)
#endif
constexpr int NMMA = MCMIL_MMA_WARPS;            // sample tc is issued by warp tc % NMMA
static_assert(NMMA == 2 || NMMA == 4, "MMA-issue warps: 2 or 4");
#ifndef MCMIL_TMEM_BUFS
#define MCMIL_TMEM_BUFS 3   // accumulator buffers in TMEM (3 x 160 of the 512 columns)
#endif
constexpr int NBUF = MCMIL_TMEM_BUFS;            // sample tc accumulates into buffer tc % NBUF
static_assert(NBUF == 2 || (NBUF == 3 && NMMA == 4), "accumulator buffers: 2, or 3 with four issue warps");
constexpr int PRODUCER_WARP0 = 4, MMA_WARP = PRODUCER_WARP0 + TEAMS * TEAM_WARPS, LOAD_WARP = MMA_WARP + 1;
constexpr int TC_THREADS = (MMA_WARP + NMMA) * 32;

constexpr uint32_t SM_W = 0;
constexpr uint32_t SM_RING = SM_W + NSLICE * SLICE_BYTES_W;        // 139264: RING_SLOTS x 8 KB
#ifndef MCMIL_RING_EXTRA
#define MCMIL_RING_EXTRA 0   // extra ring slots per producer team beyond one sample (1 = the 2 x 8 KB of smem left over;
                             // measured: no gain once the producers stopped waiting, profiles/r1_experiments.md)
#endif
constexpr int RING_EXTRA = MCMIL_RING_EXTRA;
constexpr int TEAM_SLOTS = TEAM_SLICES + RING_EXTRA;       // ring slots owned by one team; its j-th slice uses slot j % TEAM_SLOTS
constexpr int RING_SLOTS = TEAMS * TEAM_SLOTS;             // slot index = TEAMS * (j % TEAM_SLOTS) + team
static_assert(RING_EXTRA == 0 || RING_EXTRA == 1, "ring depth: one sample, or one sample + one slice per team");
constexpr uint32_t SM_XCH = SM_RING + RING_SLOTS * SLICE_BYTES_A;  // [2][64][2 x 4] floats (partial sums | logit multipliers)
constexpr uint32_t SM_BAR = SM_XCH + 2 * HALF_ROWS * 2 * MAXC * 4;
constexpr uint32_t SM_TOTAL = SM_BAR + 1024;                       // 209920 (226304 with the 10-slot ring)
static_assert(SM_TOTAL <= 232448, "227 KB of dynamic shared memory per CTA");

// where the feature keep-masks come from
enum : int {
  MASK_PHILOX = 0,         // drawn in the kernel
  MASK_INJECTED = 1        // caller-provided bits (parity tests with the reference's own masks)
};

// barrier slots (8 bytes each) inside SM_BAR.  The ring barriers and the accumulator-empty barrier
// exist once per MMA-issue warp: samples are issued round-robin by NMMA warps, and a parity wait is
// only unambiguous for a waiter that sees consecutive phases of the barrier it waits on.
enum : uint32_t {
  B_FULL = 0,                         // [NMMA][8] leader only: masked A slice of both CTAs is in smem (count 2 x TEAM_WARPS)
  B_EMPTY = B_FULL + NMMA * NSLICE,   // [NMMA][8] per CTA: the MMAs reading the slot have retired     (count 1)
  B_TFULL = B_EMPTY + NMMA * NSLICE,  // [NBUF] per CTA: accumulator buffer complete                   (count 1)
  B_TEMPTY = B_TFULL + NBUF,          // [NMMA] leader only: both CTAs' epilogues drained the buffer   (count 8)
  B_WLOC = B_TEMPTY + NMMA,           // [8] per CTA: the bulk copy of W K-slice s landed              (tx)
  B_WREADY = B_WLOC + NSLICE,         // [8] leader only: both CTAs hold their halves of W K-slice s   (count 2)
  B_COUNT = B_WREADY + NSLICE,
  TMEM_SLOT = 100                     // uint32 at SM_BAR + 8*100
};
static_assert(B_COUNT <= TMEM_SLOT && 8 * (TMEM_SLOT + 1) <= 1024, "barrier area");

// TMEM columns of one accumulator buffer (2x2 layout of the pair MMAs: lanes 0..63 hold the
// columns fed by the leader CTA's W rows, lanes 64..127 those of the peer CTA's W rows):
//   [0,72)   MMA "a" (N=144): 32 tanh | 32 sigmoid | 8 score columns
//   [96,160) MMA "b" (N=128): 32 tanh | 32 sigmoid of hidden units 64..127
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_A = 0, TM_B = 96, TM_BUF_STRIDE = 160;
static_assert(NBUF * TM_BUF_STRIDE <= TMEM_COLS, "TMEM columns");

struct ProjParams {
  const void* H;           // [R][512] packed features, fp32 or (h_f16) fp16
  int h_f16;
  const uint8_t* wmain;    // this set: [2][8][17 KB]
  const TileDesc* tiles;
  float* logits;           // [T][C][Rp]
  float* scores;           // [T][C][Rp]
  const uint32_t* inj_feat;  // [T][R][16] or null
  const uint32_t* inj_attn;  // [T][C][Rw] or null (bit = packed row)
  float* dbg;              // optional raw accumulator dump of each pair's first (tile, t)
  int n_tiles, T, C, R, Rp, Rw;   // Rp: plane columns (row stride of logits / scores); Rw: words per row of inj_attn
  int n_out;               // heads produced by this launch (shared: C, separate: 1)
  int head0;               // first head index written by this launch
  int t_offset, bag_offset;
  uint32_t thr_f, thr_a;
  float sf, hsf, sa;       // 1/(1-p_f), 0.5/(1-p_f), 1/(1-p_a)
  PhiloxKey key;
  EpiConst epi;
};

__device__ __forceinline__ uint32_t bar_addr(uint32_t sbase, uint32_t slot) { return sbase + SM_BAR + slot * 8; }

// -DMCMIL_EXP_TRACE: the warps of one leader CTA time-stamp their barrier events for samples [64, 68)
#ifdef MCMIL_EXP_TRACE
#define TRACE_DECL long long trc[64]; for (int i_ = 0; i_ < 64; ++i_) trc[i_] = 0;
#define TRACE(tcv, slot) do { if ((tcv) >= 64u && (tcv) < 68u) trc[((tcv) - 64u) * 16 + (slot)] = clock64() - k_t0; } while (0)
#define TRACE_DUMP(role) do { if (pair == 3 && rank == 0 && lane == 0) { for (int a_ = 0; a_ < 64; ++a_) \
    if (trc[a_] != 0) printf("TR %s %d %d %d %lld\n", role, warp, 64 + a_ / 16, a_ % 16, trc[a_]); } } while (0)
#else
#define TRACE_DECL
#define TRACE(tcv, slot)
#define TRACE_DUMP(role)
#endif
// -DMCMIL_EXP_WAITSTATS: every warp of the first cluster reports the cycles it spent in barrier waits
#ifndef MCMIL_RELAXED_NS_TEMPTY
#define MCMIL_RELAXED_NS_TEMPTY 256     // issue warps waiting for the accumulator buffer (idle ~2 samples out of 4)
#endif
#ifndef MCMIL_RELAXED_NS_TFULL
#define MCMIL_RELAXED_NS_TFULL 128      // epilogue warps waiting for the next accumulator (multi-buffered)
#endif
#ifndef MCMIL_RELAXED_NS_FULL
#define MCMIL_RELAXED_NS_FULL 0         // issue warps waiting for the next K-slice: plain try_wait (32 / 64 / 128 ns polling: no difference)
#endif
#ifdef MCMIL_EXP_WAITSTATS
#define WAIT_T(acc, bar, parity) do { const long long w0_ = clock64(); mbar_wait(bar, parity); acc += clock64() - w0_; } while (0)
#define WAIT_R(acc, bar, parity, ns) do { const long long w0_ = clock64(); if ((ns) > 0) mbar_wait_relaxed(bar, parity, ns); else mbar_wait(bar, parity); acc += clock64() - w0_; } while (0)
#else
#define WAIT_T(acc, bar, parity) mbar_wait(bar, parity)
#define WAIT_R(acc, bar, parity, ns) do { if ((ns) > 0) mbar_wait_relaxed(bar, parity, ns); else mbar_wait(bar, parity); } while (0)
#endif

// keep -> all-ones half lanes.  r holds two 16-bit uniform lanes; an element is kept iff
// (lane & 0x7fff) >= thr.  Positive fp16 bit patterns order like their integer values and the
// NaN patterns (> 0x7c00) must count as "large": compare |r| >=(unordered) thr as fp16x2.
__device__ __forceinline__ uint32_t keep_mask2(uint32_t r, uint32_t thr2) {
  uint32_t a, m;
  asm("abs.f16x2 %0, %1;" : "=r"(a) : "r"(r));
  asm("set.geu.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(thr2));
  return m;
}
// Same predicate with ALU-pipe integer ops only (HSET2 shares the fmaheavy pipe with the Philox
// wide multiplies): set bit 15 of both lanes, subtract the thresholds (no borrow can cross lanes),
// replicate each lane's bit 15 over the lane with one PRMT.
__device__ __forceinline__ uint32_t keep_mask2_alu(uint32_t r, uint32_t thr2) {
  const uint32_t b = (r | 0x80008000u) - thr2;
  uint32_t m;
  asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(m) : "r"(b));
  return m;
}
// The four lane-pair keep masks of one (row, chunk) from its 8 primary bytes (wa: features 0..3,
// wb: 4..7) and the refinement byte `si` of rw (philox.cuh): each 16-bit lane is built as
// (primary << 8) | refinement by one PRMT and compared by one HSET2.
__device__ __forceinline__ uint4 keep_masks(uint32_t wa, uint32_t wb, uint32_t rw, int si, uint32_t thr2) {
  const uint32_t r = 4u + (uint32_t)si;
  const uint32_t s01 = 0x1000u | (r << 8) | r, s23 = 0x3020u | (r << 8) | r;     // [R, b0, R, b1], [R, b2, R, b3]
  return make_uint4(keep_mask2(__byte_perm(wa, rw, s01), thr2), keep_mask2(__byte_perm(wa, rw, s23), thr2),
                    keep_mask2(__byte_perm(wb, rw, s01), thr2), keep_mask2(__byte_perm(wb, rw, s23), thr2));
}
// The same keep decision with ALU-pipe integer operations only (no HSET2: that shares the fmaheavy pipe with the
// Philox wide multiplies, the busiest pipe of the kernel).  With thr = T7 * 256 + T8 (T7 = top 7 bits):
//   keep  <=>  (P7 << 8 | R8) >= thr  <=>  P7 + [R8 >= T8] >= T7 + 1        (P7 = primary byte & 0x7f)
// so for the four primary bytes of a word at once:  w' = (w | 0x80808080) - (T7 + 1 - c) * 0x01010101, c = [R8 >= T8];
// every byte stays in [0, 255] (no borrow crosses bytes) and its top bit is the keep flag, which one PRMT with sign
// replication turns into the 16-bit lane masks.  The flags c of the 16 (row slot, slice) chunks of a sample come
// from the refinement words the same way, on 16-bit fields: byte 1 / byte 3 of fx = c of this team's slices 0 / 2,
// of fy = c of slices 1 / 3.  Bit-identical to keep_masks() for every threshold (also thr > 0x7c00, which the fp16
// compare cannot represent: no p_f restriction on this path).
__device__ __forceinline__ void ref_flags(uint32_t rw, uint32_t t8x2, uint32_t& fx, uint32_t& fy) {
  fx = ((rw & 0x00FF00FFu) | 0x01000100u) - t8x2;
  fy = (((rw >> 8) & 0x00FF00FFu) | 0x01000100u) - t8x2;
}
template <int SI>
__device__ __forceinline__ uint4 keep_masks_int(uint32_t wa, uint32_t wb, uint32_t fx, uint32_t fy, uint32_t neg_d) {
  const uint32_t cc = __byte_perm((SI & 1) ? fy : fx, 0u, (SI & 2) ? 0x3333u : 0x1111u);      // c * 0x01010101
  const uint32_t a = (wa | 0x80808080u) + neg_d + cc, b = (wb | 0x80808080u) + neg_d + cc;
  uint4 m;
  asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(m.x) : "r"(a));       // features 0, 1: top bit of byte 0 / 1 over a 16-bit lane
  asm("prmt.b32 %0, %1, 0, 0xBBAA;" : "=r"(m.y) : "r"(a));       // features 2, 3
  asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(m.z) : "r"(b));       // features 4, 5
  asm("prmt.b32 %0, %1, 0, 0xBBAA;" : "=r"(m.w) : "r"(b));       // features 6, 7
  return m;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One TMEM lane half of the gate epilogue.  Hidden units are processed in pairs (d, d+1) with packed fp32x2
// FMAs (FFMA2): pre-activations, the gate product and the w_c dot products each take one instruction per pair,
// 3 + NOUT FMA-type + 4 MUFU instructions per pair instead of 2 x (3 + NOUT) + 4.  acc2[c] holds the
// (even d, odd d) partial sums of head c.
template <int HALF, int NOUT, bool DEBUG>
__device__ __forceinline__ void epilogue_half(const ProjParams& P, uint32_t tbuf, float (&acc)[MAXC],
                                              float* dbg_row) {
  uint64_t acc2[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc2[c] = 0ull;
  const uint64_t sf2 = pack2(P.sf, P.sf), hsf2 = pack2(P.hsf, P.hsf);
#pragma unroll
  for (int part = 0; part < 2; ++part) {            // MMA "a" columns, then MMA "b" columns
    const uint32_t tcol = tbuf + (part == 0 ? TM_A : TM_B);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t v[16], u[16];
      tmem_ld16(tcol + 16 * j, v);
      tmem_ld16(tcol + 32 + 16 * j, u);
      tmem_ld_wait();
      if constexpr (DEBUG) {
        if (dbg_row) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            dbg_row[64 * part + 16 * j + i] = __uint_as_float(v[i]);
            dbg_row[64 * part + 32 + 16 * j + i] = __uint_as_float(u[i]);
          }
        }
      }
#ifdef MCMIL_EXP_NO_EPI_MATH
      acc[0] += __uint_as_float(v[0] ^ u[0]);
      continue;
#endif
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const int pr = (64 * part + 32 * HALF + 16 * j + i) >> 1;         // pair of hidden units (d, d+1)
        // one 16-byte constant load: (bv[d], bv[d+1]) | (0.5 bu[d], 0.5 bu[d+1]) as two aligned 64-bit operands
        const ulonglong2 cb = *reinterpret_cast<const ulonglong2*>(&P.epi.vb[pr]);
        const uint64_t xv = fma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sf2, cb.x);
        const uint64_t xu = fma2(pack2(__uint_as_float(u[i]), __uint_as_float(u[i + 1])), hsf2, cb.y);
        float xv0, xv1, xu0, xu1;
        unpack2(xv, xv0, xv1);
        unpack2(xu, xu0, xu1);
        const uint64_t av = pack2(tanh_approx(xv0), tanh_approx(xv1));
        const uint64_t au = pack2(tanh_approx(xu0), tanh_approx(xu1));
        const uint64_t g2 = fma2(av, au, av);          // 2 * tanh(.) * sigmoid(.)
        const ulonglong2 h01 = *reinterpret_cast<const ulonglong2*>(&P.epi.hw[pr][0]);
        acc2[0] = fma2(g2, h01.x, acc2[0]);
        if constexpr (NOUT > 1) acc2[1] = fma2(g2, h01.y, acc2[1]);
        if constexpr (NOUT > 2) {
          const ulonglong2 h23 = *reinterpret_cast<const ulonglong2*>(&P.epi.hw[pr][2]);
          acc2[2] = fma2(g2, h23.x, acc2[2]);
          if constexpr (NOUT > 3) acc2[3] = fma2(g2, h23.y, acc2[3]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NOUT; ++c) {
    float lo, hi;
    unpack2(acc2[c], lo, hi);
    acc[c] += lo + hi;
  }
}

template <int NOUT, int MASK, bool DEBUG, int ROUNDS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
proj_tc_kernel(const __grid_constant__ ProjParams P) {
  constexpr bool INJECT = MASK == MASK_INJECTED;            // logit masks injected too
  constexpr bool DRAW = MASK == MASK_PHILOX;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  long long wait_a = 0, wait_b = 0;          // MCMIL_EXP_WAITSTATS only
  const long long k_t0 = clock64();

  // contiguous slice of the (tile, t) unit space owned by this pair
  const long long U = (long long)P.n_tiles * P.T;
  const long long u_begin = U * pair / n_pairs, u_end = U * (pair + 1) / n_pairs;

  if ((sbase & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NMMA * NSLICE; ++s) {
      mbar_init(bar_addr(sbase, B_FULL + s), 2 * TEAM_WARPS);
      mbar_init(bar_addr(sbase, B_EMPTY + s), 1);
    }
    for (int b = 0; b < NBUF; ++b) mbar_init(bar_addr(sbase, B_TFULL + b), 1);
    for (int b = 0; b < NMMA; ++b) mbar_init(bar_addr(sbase, B_TEMPTY + b), 8);
    for (int s = 0; s < NSLICE; ++s) {
      mbar_init(bar_addr(sbase, B_WLOC + s), 1);
      mbar_init(bar_addr(sbase, B_WREADY + s), 2);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc_cg2(sbase + SM_BAR + 8 * TMEM_SLOT, TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  // TMA loader: this CTA's W rows, once per kernel, one barrier per K-slice so the first MMAs can start as soon
  // as slice 0 has landed (a single-bag call is ~11 samples per pair: the 136 KB load is not negligible there).
  // Issued before the cluster barrier (only this CTA's own smem and barriers are involved) and before
  // griddepcontrol.wait: INVARIANT — the W image is written once by mcmil_weights_create, which synchronises its
  // stream before returning, so no kernel that precedes this one in any stream can still be writing it.
  if (warp == LOAD_WARP && lane == 0) {
    for (int s = 0; s < NSLICE; ++s) {
      const uint32_t wloc = bar_addr(sbase, B_WLOC + s);
      mbar_expect_tx(wloc, SLICE_BYTES_W);
      bulk_g2s(sbase + SM_W + s * SLICE_BYTES_W, P.wmain + (size_t)(rank * NSLICE + s) * SLICE_BYTES_W, SLICE_BYTES_W, wloc);
    }
  }
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * TMEM_SLOT);

  // PDL: everything above and the W load below may overlap the tail of the previous kernel in the
  // stream; the features (producers) and the logits / scores planes (epilogue; still read by the
  // previous call's reduction kernels) may only be touched after griddepcontrol.wait.
  grid_dep_launch();
  // Register file: 16 warps x 128 registers at launch.  The four issue warps (one warpgroup) need ~40 and hand
  // the rest to the two producer warpgroups, whose 64-register feature tile + Philox state otherwise forces
  // ptxas to rematerialise addresses and spill inside the K-slice loop.
#ifndef MCMIL_NO_SETMAXNREG
  constexpr int REGS_MMA = 56, REGS_PRODUCER = 160;
  static_assert(128 + 2 * REGS_PRODUCER + REGS_MMA <= 512, "per sub-partition: 1 epilogue + 2 producer + 1 issue warp");
#endif
  if (warp >= MMA_WARP) {
#ifndef MCMIL_NO_SETMAXNREG
    reg_dealloc<REGS_MMA>();
#endif
    // ------------------------------------------------------------ W slices landed -> tell the leader (after the
    // cluster barrier: the leader's barriers are initialised)
    if (warp == LOAD_WARP && lane == 0) {
      for (int s = 0; s < NSLICE; ++s) {
        mbar_wait(bar_addr(sbase, B_WLOC + s), 0);
        mbar_arrive_cluster(mapa(bar_addr(sbase, B_WREADY + s), 0));
      }
    }
    __syncwarp();
    // ------------------------------------------------------------ MMA issuers (leader CTA)
    // NMMA issue warps take the samples round-robin: a single warp shares its scheduler with three
    // busy warps and needs ~500 cycles of issue latency per K-slice, twice the 272 cycles of tensor
    // work it launches (profiles/r1).  Each warp runs its loop warp-uniformly
    // (addresses / descriptors in uniform registers: a lane-0-only loop makes ptxas emit ELECT +
    // R2UR chains per MMA); one elected lane issues the tcgen05 ops.
    if (rank == 0) {
#ifdef MCMIL_EXP_NSPLIT      // timing experiment only (results are garbage): other N splits of the two MMAs per K-step
      constexpr uint32_t IDESC_A = umma_idesc_f16(128, MCMIL_EXP_NSPLIT);
      constexpr uint32_t IDESC_B = umma_idesc_f16(128, 272 - MCMIL_EXP_NSPLIT);
#else
      constexpr uint32_t IDESC_A = umma_idesc_f16(128, 2 * W_ROWS_A);
      constexpr uint32_t IDESC_B = umma_idesc_f16(128, 2 * W_ROWS_B);
#endif
      const uint32_t q = (uint32_t)(warp - MMA_WARP);          // this warp issues the samples tc = q (mod NMMA)
      uint32_t mbuf = q % NBUF;                                 // ... into TMEM accumulator buffer tc % NBUF
      uint32_t mslot = (q * TEAM_SLICES) % TEAM_SLOTS;          // ring slot (per team) of the sample's first slice
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t adesc0 = umma_desc_sw128(sbase + SM_RING);
      const uint64_t bdesc0 = umma_desc_sw128(sbase + SM_W);
      uint32_t j = 0;                                           // this warp's sample counter
      TRACE_DECL
#ifdef MCMIL_EXP_PRODUCER_ONLY
      if (true) goto exp_skip_mma;
#endif
      for (long long u = u_begin + q; u < u_end; u += NMMA, ++j) {
        TRACE(NMMA * j + q, 0);
        // buffer tc % NBUF was last used by sample tc - NBUF, whose epilogue arrives on TEMPTY[(tc - NBUF) % NMMA]:
        // that is phase j of TEMPTY[q - NBUF] for q >= NBUF, phase j - 1 of TEMPTY[q + NMMA - NBUF] otherwise
        const uint32_t da = tmem_u + mbuf * TM_BUF_STRIDE + TM_A, db = tmem_u + mbuf * TM_BUF_STRIDE + TM_B;
        WAIT_R(wait_b, bar_addr(sbase, B_TEMPTY + ((q + NMMA - NBUF) & (NMMA - 1))), q >= NBUF ? (j & 1) : ((j & 1) ^ 1), MCMIL_RELAXED_NS_TEMPTY);
        TRACE(NMMA * j + q, 1);
        tc_fence_after();
        bool full_ready = false;                                // early probe of the next slice's FULL barrier
#pragma unroll 1
        for (int s = 0; s < NSLICE; ++s) {
          if (j == 0) mbar_wait(bar_addr(sbase, B_WREADY + s), 0);     // this warp's first sample: W slice s of both CTAs
          if (!full_ready) WAIT_R(wait_a, bar_addr(sbase, B_FULL + q * NSLICE + s), j & 1, MCMIL_RELAXED_NS_FULL);
          tc_fence_after();
#ifndef MCMIL_NO_EARLY_PROBE
          full_ready = s + 1 < NSLICE && mbar_test_wait(bar_addr(sbase, B_FULL + q * NSLICE + s + 1), j & 1);
#endif
          if (elect_one()) {
            // start-address field is (addr >> 4): advancing by bytes/16 stays inside the 14-bit field
            uint32_t m = mslot + (uint32_t)(s / TEAMS);
            if (m >= TEAM_SLOTS) m -= TEAM_SLOTS;
            const uint64_t ad = adesc0 + (uint64_t)((m * TEAMS + (uint32_t)(s % TEAMS)) * (SLICE_BYTES_A >> 4));
            const uint64_t bd = bdesc0 + (uint64_t)(s * (SLICE_BYTES_W >> 4));
#if defined(MCMIL_EXP_MMA_ORDER)
#pragma unroll
            for (int kk = 0; kk < KSLICE / 16; ++kk) umma_f16_cg2(da, ad + 2 * kk, bd + 2 * kk, IDESC_A, (s | kk) != 0);
#pragma unroll
            for (int kk = 0; kk < KSLICE / 16; ++kk)
              umma_f16_cg2(db, ad + 2 * kk, bd + ((W_ROWS_A * 128) >> 4) + 2 * kk, IDESC_B, (s | kk) != 0);
#elif defined(MCMIL_EXP_MMA_A_ONLY)
#pragma unroll
            for (int kk = 0; kk < KSLICE / 16; ++kk) umma_f16_cg2(da, ad + 2 * kk, bd + 2 * kk, IDESC_A, (s | kk) != 0);
#else
#pragma unroll
            for (int kk = 0; kk < KSLICE / 16; ++kk) {
              umma_f16_cg2(da, ad + 2 * kk, bd + 2 * kk, IDESC_A, (s | kk) != 0);
              umma_f16_cg2(db, ad + 2 * kk, bd + ((W_ROWS_A * 128) >> 4) + 2 * kk, IDESC_B, (s | kk) != 0);
            }
#endif
            umma_commit_cg2_mc(bar_addr(sbase, B_EMPTY + q * NSLICE + s), 3);
          }
          __syncwarp();
          TRACE(NMMA * j + q, 2 + s);
        }
        if (elect_one()) umma_commit_cg2_mc(bar_addr(sbase, B_TFULL + mbuf), 3);
        __syncwarp();
        mbuf = (mbuf + NMMA) % NBUF;
        mslot = (mslot + NMMA * TEAM_SLICES) % TEAM_SLOTS;
      }
      TRACE_DUMP("mma");
#ifdef MCMIL_EXP_PRODUCER_ONLY
      exp_skip_mma:;
#endif
    }
  } else if (warp >= PRODUCER_WARP0) {
    // ------------------------------------------------------------ producers: masked A slices
    // Two teams of four warps: team 0 fills the even K-slices, team 1 the odd ones (16 rows x 8
    // chunks per warp and slice = 4 chunks per thread).  The two producer warps that share an
    // SM sub-partition belong to different teams.  The ring is a full sample deep (slot = K-slice),
    // so a warp only ever waits for the MMAs of the PREVIOUS sample.
#ifndef MCMIL_NO_SETMAXNREG
    reg_alloc<REGS_PRODUCER>();
#endif
    const int pw = warp - PRODUCER_WARP0;
    const int team = pw >> 2, wt = pw & 3;           // warps of one team sit on the four different schedulers
    // thread-constant addresses, pinned in registers (see ptx::pin)
    const uint32_t full_team = pin(mapa(bar_addr(sbase, B_FULL), 0) + (uint32_t)team * 8u);    // + (set + TEAMS*si) * 8
    const uint32_t empty_team = pin(bar_addr(sbase, B_EMPTY) + (uint32_t)team * 8u);
#ifdef MCMIL_MASK_HSET2
    const uint32_t thr2 = pin(P.thr_f | (P.thr_f << 16));
#else
    const uint32_t neg_d = pin(0u - ((P.thr_f >> 8) + 1u) * 0x01010101u);     // -(T7 + 1) in every byte
    const uint32_t t8x2 = pin((P.thr_f & 0xFFu) * 0x00010001u);
#endif
    const int chunk = lane & 7;
    const uint32_t q0 = pin((uint32_t)(team * 8 + chunk));       // Philox chunk index of this thread in slice `team`
    const uint32_t lane0 = pin(lane == 0 ? 1u : 0u);
    int rowi[4]; uint32_t off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rowi[i] = wt * 16 + i * 4 + (lane >> 3);
      off[i] = pin(sbase + SM_RING + (uint32_t)team * SLICE_BYTES_A + (uint32_t)rowi[i] * 128u +
                   (uint32_t)((chunk ^ (rowi[i] & 7)) << 4));             // ring address of slice `team`; + TEAMS*si slices
    }
    uint32_t tc = 0;                                  // samples processed so far by this pair
    bool slot_free = false;                           // early probe result for the next ring slot
    uint32_t pslot = 0;                               // ring slot (within this team's TEAM_SLOTS) of the next slice
    // The slot of this team's slice (tc, si) was last used by its slice RING_EXTRA earlier than (tc-1, si):
    // its EMPTY barrier belongs to the issue warp of that sample.
    auto empty_of = [&](uint32_t tcv, int siv, uint32_t& bar, uint32_t& parity) -> bool {
      constexpr uint32_t LOGM_ = NMMA == 4 ? 2 : 1;
      uint32_t tp = tcv - 1u;
      int sp = siv - RING_EXTRA;
      if (sp < 0) { sp += TEAM_SLICES; tp -= 1u; }
      bar = empty_team + ((tp & (NMMA - 1)) * NSLICE + TEAMS * sp) * 8;
      parity = (tp >> LOGM_) & 1u;
      return (int32_t)tp >= 0 && tcv != 0u;
    };
    TRACE_DECL
    // the tile table belongs to the plan (written once, before any launch): the first descriptor is fetched before
    // griddepcontrol.wait, off the critical path of a single-bag call
    const TileDesc td_first = P.tiles[(int)(u_begin / P.T)];
    grid_dep_wait();
    for (long long u = u_begin; u < u_end;) {
      const int ti = (int)(u / P.T);
      const int t_begin = (int)(u - (long long)ti * P.T);
      const long long u_next = (long long)(ti + 1) * P.T;
      const int t_end = (int)((u_next < u_end ? u_next : u_end) - (long long)ti * P.T);
      const TileDesc td = u == u_begin ? td_first : P.tiles[ti];
      const uint32_t bag = (uint32_t)(P.bag_offset + td.gbag);
      // this thread's 16 chunks (4 slices of its team x 4 row slots) of the fp16 feature tile: registers
      uint4 hreg[TEAM_SLICES][4];
      uint32_t nrow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int trow = (int)rank * HALF_ROWS + rowi[i];
        nrow[i] = (uint32_t)(td.n0 + trow);
#pragma unroll
        for (int si = 0; si < TEAM_SLICES; ++si) {
          uint4 packed = make_uint4(0, 0, 0, 0);
          if (trow < td.nrows) {
            const size_t e0 = (size_t)(td.row0 + trow) * L + (TEAMS * si + team) * KSLICE + chunk * 8;
            if (P.h_f16) {                                   // fp16 features: the tile is already in MMA precision
              packed = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(P.H) + e0));
            } else {
              const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(P.H) + e0);
              const float4 a = __ldg(src), b = __ldg(src + 1);
              packed.x = pack_half2(a.x, a.y); packed.y = pack_half2(a.z, a.w);
              packed.z = pack_half2(b.x, b.y); packed.w = pack_half2(b.z, b.w);
            }
          }
          hreg[si][i] = packed;
        }
      }
      // Software pipeline: the Philox words of the NEXT slice are drawn next to the masking /
      // st.shared of the CURRENT one, so wide multiplies (fmaheavy pipe) and ALU / LSU work mix.
      // Mask convention (philox.cuh): one primary call serves this thread's chunk of two rows (row
      // slots 0|1 and 2|3 differ in patch bit 2), one refinement call per sample serves its 16
      // (row slot, slice) chunks: 9 calls per sample and thread.
      static_assert(!DRAW || (TEAMS == 2 && TEAM_SLICES == 4), "refinement-byte indexing assumes two producer teams");
      uint4 rnd[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};   // primary bytes: [0] row slots 0 (x,y) | 1 (z,w), [1] slots 2 | 3
      uint4 ref = make_uint4(0, 0, 0, 0);                                // word i: row slot i, byte si: this team's si-th slice
#if !defined(MCMIL_PHILOX_BATCH2) && !defined(MCMIL_MASK_HSET2)
      uint4 q1[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)}, q2[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      uint4 ref_hold = make_uint4(0, 0, 0, 0);                           // drawn two slices ahead (see the slice loop)
#endif
#ifndef MCMIL_MASK_HSET2
      uint32_t fx[4] = {0, 0, 0, 0}, fy[4] = {0, 0, 0, 0};               // [R8 >= T8] flags of the sample's 16 chunks (ref_flags)
#endif
      if constexpr (DRAW) {
        const uint32_t tg0 = (uint32_t)(P.t_offset + t_begin);
        ref = philox4x32<ROUNDS>(REF_CHUNK_BASE + q0, nrow[0], tg0, bag, P.key);
#ifndef MCMIL_MASK_HSET2
        ref_flags(ref.x, t8x2, fx[0], fy[0]); ref_flags(ref.y, t8x2, fx[1], fy[1]);
        ref_flags(ref.z, t8x2, fx[2], fy[2]); ref_flags(ref.w, t8x2, fx[3], fy[3]);
#endif
        rnd[0] = philox4x32<ROUNDS>(q0, nrow[0], tg0, bag, P.key);
        rnd[1] = philox4x32<ROUNDS>(q0, nrow[2], tg0, bag, P.key);
      }
#pragma unroll 1
      for (int t = t_begin; t < t_end; ++t, ++tc) {
        const uint32_t tg = (uint32_t)(P.t_offset + t);
        // slot s was last read by the MMAs of sample tc-1, issued by warp (tc-1)&1 as its ((tc-1)>>1)-th
        const uint32_t full_set = (tc & (NMMA - 1)) * NSLICE;
#pragma unroll
        for (int si = 0; si < TEAM_SLICES; ++si) {
          const int s = TEAMS * si + team;
#ifndef MCMIL_EXP_PRODUCER_ONLY
          TRACE(tc, 4 * si);
          {
            uint32_t ebar, epar;
            const bool used = empty_of(tc, si, ebar, epar);
#ifndef MCMIL_NO_EARLY_PROBE
            if (used && !slot_free) WAIT_T(wait_a, ebar, epar);
#else
            if (used) WAIT_T(wait_a, ebar, epar);
#endif
          }
          TRACE(tc, 4 * si + 1);
#endif
          const uint32_t slot_off = pslot * (uint32_t)(TEAMS * SLICE_BYTES_A);
          pslot = pslot + 1 == TEAM_SLOTS ? 0u : pslot + 1;
#ifdef MCMIL_EXP_MMA_ONLY       // experiment: no producer work at all, the MMA / epilogue chain runs flat out
          __syncwarp();
          if (lane0) mbar_arrive_cluster(full_team + (full_set + TEAMS * si) * 8);
          continue;
#endif
          uint4 nxt[2], ref_nxt = ref;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 h = hreg[si][i];
            uint4 m;                                             // lane-pair keep masks of this (row slot, chunk)
            if constexpr (DRAW) {
              const uint32_t wa = (i & 1) ? rnd[i >> 1].z : rnd[i >> 1].x, wb = (i & 1) ? rnd[i >> 1].w : rnd[i >> 1].y;
#ifdef MCMIL_MASK_HSET2
              const uint32_t rw = i == 0 ? ref.x : i == 1 ? ref.y : i == 2 ? ref.z : ref.w;
#ifdef MCMIL_EXP_NO_MASK
              m = make_uint4(~0u, ~0u, ~0u, ~0u);
              asm volatile("" :: "r"(wa), "r"(wb), "r"(rw));
#else
              m = keep_masks(wa, wb, rw, si, thr2);
#endif
#else
              m = si == 0 ? keep_masks_int<0>(wa, wb, fx[i], fy[i], neg_d) : si == 1 ? keep_masks_int<1>(wa, wb, fx[i], fy[i], neg_d)
                : si == 2 ? keep_masks_int<2>(wa, wb, fx[i], fy[i], neg_d) : keep_masks_int<3>(wa, wb, fx[i], fy[i], neg_d);
#endif
            } else {
              const int trow = (int)rank * HALF_ROWS + rowi[i];
              uint32_t bits = 0;
              if (trow < td.nrows)
                bits = reinterpret_cast<const uint8_t*>(P.inj_feat)[((size_t)t * P.R + td.row0 + trow) * 64 + s * 8 + chunk];
              m.x = (bits & 1u ? 0x0000FFFFu : 0u) | (bits & 2u ? 0xFFFF0000u : 0u);
              m.y = (bits & 4u ? 0x0000FFFFu : 0u) | (bits & 8u ? 0xFFFF0000u : 0u);
              m.z = (bits & 16u ? 0x0000FFFFu : 0u) | (bits & 32u ? 0xFFFF0000u : 0u);
              m.w = (bits & 64u ? 0x0000FFFFu : 0u) | (bits & 128u ? 0xFFFF0000u : 0u);
            }
            sts128(off[i] + slot_off, make_uint4(h.x & m.x, h.y & m.y, h.z & m.z, h.w & m.w));
          }
#if !defined(MCMIL_NO_EARLY_PROBE) && !defined(MCMIL_EXP_PRODUCER_ONLY)
          // Probe the barrier of the NEXT slot now: the answer (an ~200-cycle round trip through the
          // barrier unit while the tensor core streams operands out of smem) arrives during the Philox
          // block below; the blocking wait above is only entered when the slot is really still in use.
          {
            const bool last = si == TEAM_SLICES - 1;
            uint32_t nbar, npar;
            const bool used = empty_of(last ? tc + 1u : tc, last ? 0 : si + 1, nbar, npar);
            slot_free = used && mbar_test_wait(nbar, npar);
          }
#endif
          if constexpr (DRAW) {
#if !defined(MCMIL_PHILOX_BATCH2) && !defined(MCMIL_MASK_HSET2)
            // Four (five) independent Philox chains every OTHER slice instead of two (three) every slice: with two
            // chains the LOP3 of a round issues 2-3 slots after the wide multiply it depends on and the warp stalls on
            // the fixed latency (ncu: `wait` is the top stall reason of the producer warps).  At slice 0: this
            // sample's slices 1 and 2; at slice 2: slice 3, slice 0 of the next sample and its refinement words.
            // (Measured 0.5-2 % faster in three A/B pairs, profiles/r2_experiments.md; -DMCMIL_PHILOX_BATCH2: the old order.)
            if (si == 0 || si == 2) {
              const uint32_t qa = q0 + (uint32_t)(TEAMS * (si + 1) * 8);
              const uint32_t qb = q0 + (uint32_t)(si == 0 ? TEAMS * 2 * 8 : 0);
              const uint32_t tb = si == 0 ? tg : tg + 1u;
              q1[0] = philox4x32<ROUNDS>(qa, nrow[0], tg, bag, P.key);
              q1[1] = philox4x32<ROUNDS>(qa, nrow[2], tg, bag, P.key);
              q2[0] = philox4x32<ROUNDS>(qb, nrow[0], tb, bag, P.key);
              q2[1] = philox4x32<ROUNDS>(qb, nrow[2], tb, bag, P.key);
              if (si == 2) ref_nxt = ref_hold = philox4x32<ROUNDS>(REF_CHUNK_BASE + q0, nrow[0], tb, bag, P.key);
            }
            nxt[0] = (si == 0 || si == 2) ? q1[0] : q2[0];
            nxt[1] = (si == 0 || si == 2) ? q1[1] : q2[1];
            if (si == TEAM_SLICES - 1) ref_nxt = ref_hold;
#else
            // next slice of this team: (s + TEAMS, t), or (team, t + 1) after the last one of the sample
            const uint32_t q_next = q0 + (uint32_t)(si < TEAM_SLICES - 1 ? TEAMS * (si + 1) * 8 : 0);
            const uint32_t t_next = si < TEAM_SLICES - 1 ? tg : tg + 1u;
            nxt[0] = philox4x32<ROUNDS>(q_next, nrow[0], t_next, bag, P.key);
            nxt[1] = philox4x32<ROUNDS>(q_next, nrow[2], t_next, bag, P.key);
            if (si == TEAM_SLICES - 1)
              ref_nxt = philox4x32<ROUNDS>(REF_CHUNK_BASE + q0, nrow[0], t_next, bag, P.key);
#endif
          }
          TRACE(tc, 4 * si + 2);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane0) mbar_arrive_cluster(full_team + (full_set + TEAMS * si) * 8);
          TRACE(tc, 4 * si + 3);
          if constexpr (DRAW) {
            rnd[0] = nxt[0]; rnd[1] = nxt[1];
#ifdef MCMIL_MASK_HSET2
            ref = ref_nxt;
#else
            if (si == TEAM_SLICES - 1) {             // the next sample's refinement words were just drawn
              ref_flags(ref_nxt.x, t8x2, fx[0], fy[0]); ref_flags(ref_nxt.y, t8x2, fx[1], fy[1]);
              ref_flags(ref_nxt.z, t8x2, fx[2], fy[2]); ref_flags(ref_nxt.w, t8x2, fx[3], fy[3]);
            }
#endif
          }
        }
      }
      u = u_next < u_end ? u_next : u_end;
    }
    TRACE_DUMP("prod");
  } else {
    // ------------------------------------------------------------ epilogue warps 0..3
    const int half = warp >> 1;                    // TMEM lanes 64..127: the peer CTA's W rows
    const int r = (warp & 1) * 32 + lane;          // patch row within this CTA's 64-row half tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t tempty_leader = mapa(bar_addr(sbase, B_TEMPTY), 0);
    float* xch = reinterpret_cast<float*>(smem + SM_XCH);
    uint32_t tc = 0, buf = 0, buf_phase = 0;      // accumulator buffer tc % NBUF and the parity of its use count
    TRACE_DECL
    const TileDesc td_first = P.tiles[(int)(u_begin / P.T)];     // plan-owned table: safe before griddepcontrol.wait
    grid_dep_wait();
#ifdef MCMIL_EXP_PRODUCER_ONLY
    for (long long u = u_end; u < u_end;) {
#else
    for (long long u = u_begin; u < u_end;) {
#endif
      const int ti = (int)(u / P.T);
      const int t_begin = (int)(u - (long long)ti * P.T);
      const long long u_next = (long long)(ti + 1) * P.T;
      const int t_end = (int)((u_next < u_end ? u_next : u_end) - (long long)ti * P.T);
      const TileDesc td = u == u_begin ? td_first : P.tiles[ti];
      const int trow = (int)rank * HALF_ROWS + r;
      const bool valid = trow < td.nrows;
      const int g = td.row0 + trow;
      for (int t = t_begin; t < t_end; ++t, ++tc) {
        TRACE(tc, 0);
        WAIT_R(wait_a, bar_addr(sbase, B_TFULL + buf), buf_phase, MCMIL_RELAXED_NS_TFULL);
        TRACE(tc, 1);
        tc_fence_after();
        float acc[MAXC] = {0.f, 0.f, 0.f, 0.f};
        float* dbg_row = nullptr;
        if constexpr (DEBUG) {
          if (P.dbg != nullptr && tc == 0)
            dbg_row = P.dbg + ((size_t)(pair * 2 + (int)rank) * 128 + warp * 32 + lane) * 136;
        }
        if (half == 0) epilogue_half<0, NOUT, DEBUG>(P, lane_base + buf * TM_BUF_STRIDE, acc, dbg_row);
        else           epilogue_half<1, NOUT, DEBUG>(P, lane_base + buf * TM_BUF_STRIDE, acc, dbg_row);
        uint32_t sc[8];
        tmem_ld8(lane_base + buf * TM_BUF_STRIDE + TM_A + 64, sc);
        tmem_ld_wait();
        if constexpr (DEBUG) {
          if (dbg_row) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dbg_row[128 + i] = __uint_as_float(sc[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + (tc & (NMMA - 1)) * 8);
        TRACE(tc, 2);
        if (++buf == NBUF) { buf = 0; buf_phase ^= 1; }
        // combine the two hidden-unit halves of each patch row.  The lower-half warps do the global stores;
        // the logit-dropout multipliers (one Philox call per row and sample) are drawn by the upper-half
        // warps on even samples and by the lower-half warps on odd ones: the work left after the TMEM drain
        // is split between the two warp pairs (= SM sub-partitions 0,1 and 2,3, which all host two producers).
        float* x = xch + ((tc & 1) * HALF_ROWS + r) * (2 * MAXC);
        const bool draws = (half == 1) == ((tc & 1) == 0);
        float mult[MAXC] = {0.f, 0.f, 0.f, 0.f};
        if (draws && valid) {
          const uint32_t tg = (uint32_t)(P.t_offset + t);
          uint4 rnd = make_uint4(0, 0, 0, 0);
          if constexpr (!INJECT)
            rnd = attn_words<ROUNDS>(0u, (uint32_t)(td.n0 + trow), tg, (uint32_t)(P.bag_offset + td.gbag), P.key);
#pragma unroll
          for (int c = 0; c < NOUT; ++c) {
            const int head = P.head0 + c;
            bool keep;
            if constexpr (!INJECT) keep = attn_keep_from(rnd, head, P.thr_a);
            else keep = (P.inj_attn[((size_t)t * P.C + head) * P.Rw + (g >> 5)] >> (g & 31)) & 1u;
            mult[c] = keep ? P.sa : 0.f;
          }
        }
        if (half == 1) {
          *reinterpret_cast<float4*>(x) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          if (draws) *reinterpret_cast<float4*>(x + MAXC) = make_float4(mult[0], mult[1], mult[2], mult[3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (half == 0 && valid) {
          const float4 xs = *reinterpret_cast<const float4*>(x);
          const float part[MAXC] = {xs.x, xs.y, xs.z, xs.w};
          if (!draws) {
            const float4 xm = *reinterpret_cast<const float4*>(x + MAXC);
            mult[0] = xm.x; mult[1] = xm.y; mult[2] = xm.z; mult[3] = xm.w;
          }
#pragma unroll
          for (int c = 0; c < NOUT; ++c) {
            const int head = P.head0 + c;
            const float logit = acc[c] + part[c] + P.epi.bw[c];
            const size_t o = ((size_t)t * P.C + head) * P.Rp + td.pcol0 + trow;
            P.logits[o] = mult[c] != 0.f ? logit * mult[c] : 0.f;   // a dropped logit is 0, not -inf (model.py:291,305)
            P.scores[o] = (__uint_as_float(sc[c]) + __uint_as_float(sc[4 + c])) * P.sf;
          }
        }
      }
      u = u_next < u_end ? u_next : u_end;
    }
    TRACE_DUMP("epi");
  }

#ifdef MCMIL_EXP_WAITSTATS
  if (pair == 3 && lane == 0)
    printf("WS rank %u warp %2d total %lld wait_a %lld wait_b %lld\n", rank, warp, clock64() - k_t0, wait_a, wait_b);
#else
  (void)wait_a; (void)wait_b; (void)k_t0;
#endif
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

cudaError_t launch_proj_tc(const Weights& w, const Plan& p, const MaskSpec& m, const void* H, int h_f16,
                           float* logits, float* scores, float* dbg, cudaStream_t st, int* launches) {
  using KernelFn = void (*)(ProjParams);
  // [rounds 10 / 7][mask source][heads per launch]
  static const KernelFn kernels[2][2][4] = {
      {{proj_tc_kernel<1, MASK_PHILOX, false, 10>, proj_tc_kernel<2, MASK_PHILOX, false, 10>,
        proj_tc_kernel<3, MASK_PHILOX, false, 10>, proj_tc_kernel<4, MASK_PHILOX, false, 10>},
       {proj_tc_kernel<1, MASK_INJECTED, false, 10>, proj_tc_kernel<2, MASK_INJECTED, false, 10>,
        proj_tc_kernel<3, MASK_INJECTED, false, 10>, proj_tc_kernel<4, MASK_INJECTED, false, 10>}},
      {{proj_tc_kernel<1, MASK_PHILOX, false, 7>, proj_tc_kernel<2, MASK_PHILOX, false, 7>,
        proj_tc_kernel<3, MASK_PHILOX, false, 7>, proj_tc_kernel<4, MASK_PHILOX, false, 7>},
       {proj_tc_kernel<1, MASK_INJECTED, false, 10>, proj_tc_kernel<2, MASK_INJECTED, false, 10>,
        proj_tc_kernel<3, MASK_INJECTED, false, 10>, proj_tc_kernel<4, MASK_INJECTED, false, 10>}}};
  static const KernelFn debug_kernel = proj_tc_kernel<2, MASK_PHILOX, true, 10>;   // raw-accumulator dump (tests only)
  if (dbg != nullptr && !(w.shared && w.C == 2 && m.inj_feat == nullptr && m.rounds == 10)) return cudaErrorInvalidValue;
  // The dynamic shared memory opt-in is a PER-DEVICE function attribute and the SM count a per-device property:
  // both are set / queried once per device ordinal, under a mutex (several host threads may drive several GPUs).
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  int sms = 0;
  {
    static std::mutex mu;
    static int sms_of[64] = {0};                     // 0 = this device has not been set up yet
    std::lock_guard<std::mutex> lock(mu);
    const bool cached = dev >= 0 && dev < 64 && sms_of[dev] != 0;
    if (!cached) {
      for (int r = 0; r < 2; ++r)
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 4; ++b) {
            cudaError_t e = cudaFuncSetAttribute(kernels[r][a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
            if (e != cudaSuccess) return e;
          }
      cudaError_t e = cudaFuncSetAttribute(debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
      if (e != cudaSuccess) return e;
      e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) sms_of[dev] = sms;
    } else {
      sms = sms_of[dev];
    }
  }
  const long long units = (long long)p.n_tiles * p.T;
  int n_pairs = sms / 2;
  if (p.sm_limit > 0 && p.sm_limit / 2 < n_pairs) n_pairs = p.sm_limit / 2 > 0 ? p.sm_limit / 2 : 1;
  if (units < n_pairs) n_pairs = (int)units;
  if (n_pairs < 1) return cudaSuccess;
  for (int s = 0; s < w.S; ++s) {
    ProjParams P;
    P.H = H; P.h_f16 = h_f16;
    P.wmain = w.d_wmain + (size_t)s * 2 * NSLICE * SLICE_BYTES_W;
    P.tiles = p.d_tiles;
    P.logits = logits; P.scores = scores;
    P.inj_feat = m.inj_feat; P.inj_attn = m.inj_attn;
    P.dbg = dbg;
    P.n_tiles = p.n_tiles; P.T = p.T; P.C = p.C; P.R = p.R; P.Rp = p.Rp; P.Rw = p.Rw;
    P.n_out = w.shared ? w.C : 1;
    P.head0 = w.shared ? 0 : s;
    P.t_offset = m.t_offset; P.bag_offset = m.bag_offset;
    P.thr_f = m.thr_f; P.thr_a = m.thr_a;
    P.sf = m.sf; P.hsf = 0.5f * m.sf; P.sa = m.sa;
    P.key = m.key;
    P.epi = w.epi[s];
    // separate attention: every head's launch redraws the same Philox masks (the same H_drop feeds all heads,
    // model.py:281,297-298); a keep-bit cache in HBM was measured no faster once a call costs 1/16 per element
    const int mask_mode = m.inj_feat != nullptr ? MASK_INJECTED : MASK_PHILOX;
    const KernelFn fn = dbg != nullptr ? debug_kernel : kernels[m.rounds == 7 ? 1 : 0][mask_mode][P.n_out - 1];
    {
      PdlLaunch L(dim3(2 * n_pairs), dim3(TC_THREADS), SM_TOTAL, st, p.sm_limit == 0);
      cudaError_t e = cudaLaunchKernelEx(&L.cfg, fn, P);
      if (e != cudaSuccess) return e;
    }
    if (launches) ++*launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace mcmil

// K4 (tensor-core path, CTA-PAIR variant): same algorithm, schedule and numerics as sky_shade_tc.cu, but two CTAs of a
// 2-CTA cluster (one TPC) run every GEMM as ONE tcgen05.mma.cta_group::2 of M = 256: each CTA keeps its own 128-pair tile
// (A operands in its shared memory, accumulators in its TMEM, its own prologue / epilogue warps) and streams only HALF of
// every weight stage (its N/2 rows of B).  That halves the L2 -> SMEM weight traffic per SM -- the 4 x 16 KB ring was
// latency-bound at ~85 GB/s per SM (profiles/r01_k4_phase_cycles.log) -- and halves the MMA-issue overhead per tile (one
// issuer thread drives both SMs).  Cross-CTA protocol: tcgen05.commit multicasts to the mbarriers of both CTAs; the
// epilogue / prologue warps of the peer CTA arrive remotely (mapa + mbarrier.arrive.release.cluster) on the leader's
// dependency barriers; the peer's idle MMA warp relays "my half of stage s has landed" to the leader.
//
// K4 (tensor-core path): DDF sky visibility fused with the cosine-weighted Lambertian sum on
// tcgen05 / TMEM, weights streamed by the TMA engine (cp.async.bulk -> UBLKCP).
//
// Replaces, for every (ray, light-direction) pair, NeuSkyFactoModel.compute_visibility
// (neusky/models/neusky_model.py:1624-1778), DDFModel.get_outputs (neusky/models/ddf_model.py:158-219),
// DirectionalDistanceField.get_outputs (neusky/fields/directional_distance_field.py:261-306),
// FiLMSiren (ns_reni/reni/field_components/film_siren.py:45-156) and the visibility-weighted einsum of
// RGBLambertianRendererWithVisibility (neusky/model_components/renderers.py:106-113).
//
// Numerics: fp16 operands (activations and weights), fp32 accumulation in TMEM, fp32 epilogues.
// Parity with the fp32 reference is stated separately for this path (tests/test_gpu_tc.py).
//
// One persistent CTA per SM processes tiles of 128 pairs (rows).  Per tile the whole DDF network
// (35->256 x5 LeakyReLU mapping net -> 2560 FiLM parameters; 15->256 x5 FiLM-SIREN trunk -> 1) runs as
// a chain of 128xNx16 tcgen05.mma instructions whose accumulators never leave TMEM:
//
//   TMEM (512 columns)  ACC_A = cols [0,256)   mapping layers 1,3,5 / trunk pre-activation Z_l
//                       ACC_B = cols [256,512) mapping layers 2,4 ; in the trunk phase split into
//                       FP0 = [256,384), FP1 = [384,512): double-buffered FiLM chunks [freq 64 | phase 64]
//   SMEM   ACT_M 68 KB  mapping activations / m5 (A operand, fp16, K-major no-swizzle canonical layout),
//                       K = 256 + a 16-wide "ones" block: columns 256,257 hold 1.0 so that every bias is
//                       a pair of extra weight columns (fp16 hi + lo) and the epilogues add nothing
//          ACT_H 64 KB  trunk activations h_l
//          IN_M 12 KB, IN_H 4 KB  first-layer inputs of the NEXT tile (written by the prologue warps)
//          ring  4 x 16 KB  weight stages, filled by cp.async.bulk from the pre-tiled fp16 blob (L2 resident)
//
//   warp 0      weight producer (one lane): walks the 147-stage stream once per tile
//   warp 1      MMA issuer (one lane): static schedule below, mbarrier-gated
//   warp 2      TMEM allocator
//   warps 4-11  epilogue (2 warps per TMEM lane quadrant): TMEM -> regs -> bias/LeakyReLU or
//               sin(freq*z+phase) -> fp16 -> SMEM A operand of the next MMA; last layer: 256->1 dot,
//               sigmoid, visibility, Lambertian accumulation (atomics into rgb_lin)
//   warps 12-15 prologue (thread = row): pair geometry, sphere exit point, local frame, NeRF PE,
//               16-level hash-grid gather of the NEXT tile while the current one is in the MMA chain
//
// FiLM folding done on the host (packing.pack_ddf_tc): freq' = 15 f + 30 and the trunk bias b are
// folded into the FiLM weights: sin(freq' * (z + b) + phase) = sin(freq' * z + phase'),
// phase' = phase + freq' * b (linear in m5, so it is one more row block of the same GEMM).
#include "nsk_common.cuh"
#include "tc_util.cuh"
#include <stdlib.h>
#include <type_traits>

namespace nsk {
namespace tcs2 {

using namespace nsk::tc;

constexpr int TM = 128;                      // rows (pairs) per tile
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4, PRO_WARP0 = 12;
constexpr int EPI_THREADS = 256, PRO_THREADS = 128;

// Weight stream per tile, in MMA issue order.  A stage is an [N][kps] fp16 operand tile:
//   M1: [256][32] [256][16]        (K = 35 features + bias hi/lo columns, padded to 48)
//   M2..M5: 8 x [256][32] + [256][16] (the 16-wide tail carries the bias hi/lo columns)
//   FP(l,c): 4 x [128][64] + [128][16]   rows 0..63 freq', rows 64..127 phase' of columns c*64..c*64+63
//   Z_0: [256][48] = [W_hi | W_hi | W_lo] against the input tile [x_hi | x_lo | x_hi] (fp16 hi/lo split of BOTH operands of the
//        first trunk layer: its pre-activation is multiplied by freq' ~ 30..45, and the fp16 rounding of the 15 direction features
//        and of W_0 was ~85 % of this path's visibility error; the split costs two extra K = 16 MMAs per tile)
//   Z_1..Z_4: 8 x [256][32]  (trunk biases are folded into phase')
constexpr int64_t STREAM_BYTES = (16384 + 8192) + 4ll * (8 * 16384 + 8192) + 20ll * (4 * 16384 + 4096) + 24576 + 4ll * 8 * 16384;
constexpr int TAIL_FLOATS = 256 /*w_final*/ + 4 /*b_final*/;
constexpr int64_t STREAM_BYTES_RANK = STREAM_BYTES / 2;        // each CTA of the pair streams its N/2 rows of every operand tile
constexpr int64_t BLOB_BYTES = STREAM_BYTES + (int64_t)TAIL_FLOATS * 4;   // [rank-0 half | rank-1 half | tail]
constexpr int KM = 272;                            // ACT_M K extent (256 + ones block)

// shared memory carve-up (bytes)
constexpr uint32_t OFF_ACT_M = 0;                  // [128][272] fp16
constexpr uint32_t OFF_ACT_H = TM * KM * 2;        // [128][256] fp16
constexpr uint32_t OFF_IN_M = OFF_ACT_H + 65536;   // [128][48] fp16
constexpr uint32_t OFF_IN_H = OFF_IN_M + 12288;    // [128][48] fp16: x_hi | x_lo | x_hi
constexpr uint32_t OFF_RING = OFF_IN_H + 12288;
constexpr uint32_t OFF_GEO = OFF_RING + NSTAGE * STAGE_BYTES;   // [2][128] x {term, pad} + fin[128]
constexpr uint32_t OFF_BAR = OFF_GEO + 2 * 128 * 4 + 128 * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32 * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory plan exceeds the 227 KB per-CTA limit");

// barrier indices
enum { B_WFULL = 0, B_WEMPTY = 4, B_INFULL = 8, B_INEMPTY = 9, B_ACCA = 10, B_MAPB = 11, B_FPFULL = 12, B_MACT = 14, B_FPFREE = 15, B_PFULL = 17, B_COUNT = 21 };

constexpr uint32_t TM_ACC_A = 0, TM_ACC_B = 256, TM_FP0 = 256, TM_FP1 = 384;

struct Params {
  const float* points; int64_t R;
  const float* normals; const float* wa; const float* inv_count; int S;
  const float* dirs; int Dp;
  const float* radiance; const int32_t* cam;
  const uint8_t* blob; const float2* table; const float* scalings; int log2_T;
  float radius, thr, sig_scale;
  float* rgb_lin; float* vis_out; float* ddf_out; float* term_out;
  int64_t n_pairs, n_tiles, n_tp;   // n_tp = tile pairs
  GridMode gm;                // DDF position grid: nerfstudio torch semantics (meta == nullptr) or an imported tiny-cuda-nn grid
  unsigned long long* prof;   // diagnostics: [grid][16] cycle counters (NULL = off)
  float* dbg;   // diagnostics: [10][128][256] activations of the first tile of CTA 0 (NULL = off)
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- static per-tile schedule ------------------------------------------------------------------
// One entry per GEMM op in MMA issue order.  `wait` is the mbarrier the issuer must see complete
// before the op (0xff = none), `commit0/1` the barriers that a tcgen05.commit arrives on after it.
struct Op {
  uint8_t wait, a_sel, shape, dst, commit0, commit1, pad0, pad1;
  uint32_t a_off;   // byte offset of the A operand tile in shared memory
  uint32_t d_col;   // first TMEM column of the accumulator
};

enum { A_IN_M = 0, A_ACT_M = 1, A_IN_H = 2, A_ACT_H = 3 };
enum { SH_M1 = 0, SH_MAP = 1, SH_FP = 2, SH_Z0 = 3, SH_Z = 4, SH_NONE = 5 };
enum { D_ACC_A = 0, D_ACC_B = 1, D_FP0 = 2, D_FP1 = 3 };
constexpr uint8_t NOB = 0xff;
constexpr int NUM_OPS = 32;

struct Schedule {
  Op ops[NUM_OPS];
  constexpr Schedule() : ops{} {
    int n = 0;
    ops[n++] = Op{B_INFULL, A_IN_M, SH_M1, D_ACC_A, B_ACCA, NOB, 0, 0, 0, 0};
    for (int i = 2; i <= 5; ++i)
      ops[n++] = Op{B_MACT, A_ACT_M, SH_MAP, (uint8_t)((i & 1) ? D_ACC_A : D_ACC_B), (uint8_t)((i & 1) ? B_ACCA : B_MAPB), NOB, 0, 0, 0, 0};
    ops[n++] = Op{B_MACT, A_ACT_M, SH_FP, D_FP0, B_FPFULL + 0, NOB, 0, 0, 0, 0};      // FP(0,0): m5 ready
    ops[n++] = Op{NOB, A_IN_H, SH_Z0, D_ACC_A, B_ACCA, B_INEMPTY, 0, 0, 0, 0};        // Z_0 ; IN buffers consumed
    ops[n++] = Op{NOB, A_ACT_M, SH_FP, D_FP1, B_FPFULL + 1, NOB, 0, 0, 0, 0};         // FP(0,1)
    for (int l = 0; l < 5; ++l) {
      ops[n++] = Op{B_FPFREE + 0, A_ACT_M, SH_FP, D_FP0, B_FPFULL + 0, NOB, 0, 0, 0, 0};   // FP(l,2) after C(l,0)
      ops[n++] = Op{B_FPFREE + 1, A_ACT_M, SH_FP, D_FP1, B_FPFULL + 1, NOB, 0, 0, 0, 0};   // FP(l,3) after C(l,1)
      if (l < 4) {
        ops[n++] = Op{B_FPFREE + 0, A_ACT_M, SH_FP, D_FP0, B_FPFULL + 0, NOB, 0, 0, 0, 0}; // FP(l+1,0) after C(l,2)
        ops[n++] = Op{B_FPFREE + 1, A_ACT_H, SH_Z, D_ACC_A, B_ACCA, NOB, 0, 0, 0, 0};      // Z_{l+1} after C(l,3): h_l complete
        ops[n++] = Op{NOB, A_ACT_M, SH_FP, D_FP1, B_FPFULL + 1, NOB, 0, 0, 0, 0};          // FP(l+1,1)
      } else {
        ops[n++] = Op{B_FPFREE + 0, 0, SH_NONE, 0, NOB, NOB, 0, 0, 0, 0};                  // drain C(4,2)
        ops[n++] = Op{B_FPFREE + 1, 0, SH_NONE, 0, NOB, NOB, 0, 0, 0, 0};                  // drain C(4,3): TMEM free
      }
    }
    for (int i = 0; i < NUM_OPS; ++i) {
      ops[i].a_off = ops[i].a_sel == A_IN_M ? OFF_IN_M : ops[i].a_sel == A_ACT_M ? OFF_ACT_M : ops[i].a_sel == A_IN_H ? OFF_IN_H : OFF_ACT_H;
      ops[i].d_col = ops[i].dst == D_ACC_A ? TM_ACC_A : ops[i].dst == D_ACC_B ? TM_ACC_B : ops[i].dst == D_FP0 ? TM_FP0 : TM_FP1;
    }
  }
};
__constant__ Schedule c_sched = Schedule();
// shape -> (N, number of full stages, K columns per full stage, K columns of the tail stage)
__constant__ int c_shape_N[6] = {256, 256, 128, 256, 256, 0};
__constant__ int c_shape_nfull[6] = {0, 4, 2, 0, 4, 0};
__constant__ int c_shape_kps[6] = {64, 64, 128, 64, 64, 0};
__constant__ int c_shape_ktail[6] = {48, 16, 16, 48, 0, 0};
// The same tables as compile-time constants: the MMA issuer is ONE thread and the serial resource of the kernel, so its walk over
// the schedule is unrolled with every operand an immediate (no __constant__ lookups, no runtime shape arithmetic between MMAs).
constexpr Schedule k_sched = Schedule();
constexpr int k_shape_N[6] = {256, 256, 128, 256, 256, 0};
constexpr int k_shape_nfull[6] = {0, 4, 2, 0, 4, 0};
constexpr int k_shape_kps[6] = {64, 64, 128, 64, 64, 0};
constexpr int k_shape_ktail[6] = {48, 16, 16, 48, 0, 0};
template <class F, int... Os>
__device__ __forceinline__ void for_each_op(F&& f, std::integer_sequence<int, Os...>) {
  (f(std::integral_constant<int, Os>{}), ...);
}

template <int VARIANT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) sky_shade_tc2_kernel(const Params P) {
  // VARIANT bit0: epilogue math stripped (diagnostic), bit1: MMAs not issued (diagnostic)
  constexpr bool PROF = (VARIANT & 4) != 0;      // per-role cycle accounting (diagnostic instantiation only)
  constexpr bool DBG = (VARIANT & 8) != 0;       // activation dump of the first tile (diagnostic instantiation only: keeps the stores out of the hot epilogue loops)
#define NSK_CLK() (PROF ? clock64() : 0ll)
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32 * 8);
  float* geo_term = reinterpret_cast<float*>(smem + OFF_GEO);                 // [2][128]
  float* fin_part = reinterpret_cast<float*>(smem + OFF_GEO + 2 * 128 * 4);  // [128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tailw = reinterpret_cast<const float*>(P.blob + STREAM_BYTES);  // w_final[256], b_final
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t tp0 = blockIdx.x >> 1, tp_step = gridDim.x >> 1;
  // arrive once per warp on a dependency barrier that lives in the LEADER CTA (local or remote)
  auto arrive_leader = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0) { if (leader) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); }
  };

  if (threadIdx.x == 0) {
    // WFULL of the LEADER counts two arrivals per use: its own producer's expect_tx and the peer's "my half has landed" relay, so the
    // issuer polls ONE barrier per weight stage (147 stages per tile)
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bars + 8 * (B_WFULL + i), leader ? 2 : 1); mbar_init(bars + 8 * (B_WEMPTY + i), 1); }
    mbar_init(bars + 8 * B_INFULL, 2 * (PRO_THREADS / 32));          // one arrival per prologue warp of both CTAs
    for (int i = 0; i < NSTAGE; ++i) mbar_init(bars + 8 * (B_PFULL + i), 1);
    mbar_init(bars + 8 * B_INEMPTY, 1);
    mbar_init(bars + 8 * B_ACCA, 1);
    mbar_init(bars + 8 * B_MAPB, 1);
    mbar_init(bars + 8 * (B_FPFULL + 0), 1); mbar_init(bars + 8 * (B_FPFULL + 1), 1);
    mbar_init(bars + 8 * B_MACT, 2 * (EPI_THREADS / 32));
    mbar_init(bars + 8 * (B_FPFREE + 0), 2 * (EPI_THREADS / 32)); mbar_init(bars + 8 * (B_FPFREE + 1), 2 * (EPI_THREADS / 32));
    fence_barrier_init();
  }
  // the "ones" block of ACT_M (k = 256..271): columns 256 and 257 are 1.0, written once
  if (threadIdx.x < TM) {
    uint8_t* d = smem + OFF_ACT_M + (uint32_t)(256 / 8) * (TM * 16) + threadIdx.x * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // fp16 1.0, 1.0, 0...
    *reinterpret_cast<uint4*>(d + TM * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 2) tmem_alloc2<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_last();
      uint32_t st = 0, ph = 0;
      long long t_wait = 0;
      const long long t_begin = NSK_CLK();
      for (int64_t tp = tp0; tp < P.n_tp; tp += tp_step) {
        const uint8_t* src = P.blob + (int64_t)rank * STREAM_BYTES_RANK;
#pragma unroll 1
        for (int o = 0; o < NUM_OPS; ++o) {
          const int sh = c_sched.ops[o].shape;
          const int N = c_shape_N[sh], nfull = c_shape_nfull[sh], kps = c_shape_kps[sh], ktail = c_shape_ktail[sh];
          const int nst = nfull + (ktail ? 1 : 0);
#pragma unroll 1
          for (int sg = 0; sg < nst; ++sg) {
            const uint32_t bytes = (uint32_t)(N >> 1) * (uint32_t)(sg < nfull ? kps : ktail) * 2u;
            mbar_wait(bars + 8 * (B_WEMPTY + st), ph ^ 1);
            mbar_arrive_expect_tx(bars + 8 * (B_WFULL + st), bytes);
            bulk_g2s_hint(sbase + OFF_RING + st * STAGE_BYTES, src, bytes, bars + 8 * (B_WFULL + st), pol);
            src += bytes;
            if (++st == NSTAGE) { st = 0; ph ^= 1; }
          }
        }
      }
      if (P.prof) { P.prof[blockIdx.x * 16 + 0] = NSK_CLK() - t_begin; P.prof[blockIdx.x * 16 + 1] = t_wait; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp walks the schedule (warp-uniform control flow); one elected lane issues the
    // tcgen05.mma / tcgen05.commit instructions.  Descriptors are advanced by adding to their low words.
    uint32_t wst = 0, wph = 0, phases = 0;
    long long t_dep = 0, t_w = 0, t_dep_map = 0, t_iss_map = 0, t_iss_fp = 0, t_iss_z = 0;
    const long long t_begin = NSK_CLK();
    const uint64_t desc_hi = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32);   // version 1, SBO = 128 B
    if (!leader) {
      // ---- peer CTA: relay "my half of stage s has landed" to the leader's PFULL barriers -----------------------
      for (int64_t tp = tp0; tp < P.n_tp; tp += tp_step) {
#pragma unroll 1
        for (int o = 0; o < NUM_OPS; ++o) {
          const int sh = c_sched.ops[o].shape;
          if (sh == SH_NONE) continue;
          const int nst = c_shape_nfull[sh] + (c_shape_ktail[sh] ? 1 : 0);
#pragma unroll 1
          for (int sg = 0; sg < nst; ++sg) {
            mbar_wait(bars + 8 * (B_WFULL + wst), wph);
            if (lane == 0) mbar_arrive_remote(bars + 8 * (B_WFULL + wst), 0);
            __syncwarp();
            if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
          }
        }
      }
    } else if (lane == 0) {
      // ---- leader CTA: one thread issues every MMA of the pair ---------------------------------------------------
      auto issue_op = [&](auto o_tag) {
        constexpr int O = decltype(o_tag)::value;
        constexpr Op op = k_sched.ops[O];
        if constexpr (op.wait != NOB) {
          const long long c0 = NSK_CLK();
          mbar_wait(bars + 8 * op.wait, (phases >> op.wait) & 1u);
          if (PROF) {
            const long long dt = NSK_CLK() - c0;
            t_dep += dt;
            if (O < 6) t_dep_map += dt;
          }
          phases ^= 1u << op.wait;
          tc_fence_after();
        }
        constexpr int sh = op.shape;
        if constexpr (sh != SH_NONE) {
          constexpr int N = k_shape_N[sh], nfull = k_shape_nfull[sh], kps = k_shape_kps[sh], ktail = k_shape_ktail[sh];
          constexpr int Nh = N >> 1;                                  // B rows held by each CTA
          constexpr uint32_t idesc = make_idesc_f16(2 * TM, N);       // M = 256 across the pair
          const uint32_t tmem_d = tmem + op.d_col;
          // A descriptor: LBO = 128 rows * 16 B; advancing one K=16 step moves 2 chunks = 4096 B
          uint64_t ad = desc_hi | ((uint64_t)((TM * 16) >> 4) << 16) | (uint64_t)(((sbase + op.a_off) >> 4) & 0x3FFF);
          constexpr uint64_t bd0 = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)((Nh * 16) >> 4) << 16);
          constexpr uint32_t b_step = (uint32_t)(2 * Nh * 16) >> 4;
          auto stage = [&](auto nmma_tag, const uint32_t first_acc) {
            constexpr int NMMA = decltype(nmma_tag)::value;
            {
              const long long c0 = NSK_CLK();
              mbar_wait(bars + 8 * (B_WFULL + wst), wph);                  // my half of the stage AND the peer's (relayed arrival)
              if (PROF) t_w += NSK_CLK() - c0;
            }
            tc_fence_after();
            const long long ci0 = NSK_CLK();
            uint64_t bd = bd0 | (uint64_t)(((sbase + OFF_RING + wst * STAGE_BYTES) >> 4) & 0x3FFF);
#pragma unroll
            for (int j = 0; j < NMMA; ++j) {
              umma_ss2(tmem_d, ad, bd, idesc, j == 0 ? first_acc : 1u);
              ad += 256;                                                   // one K=16 step = 2 chunks = 4096 B >> 4
              bd += b_step;
            }
            umma_commit2(bars + 8 * (B_WEMPTY + wst));   // stage reusable in BOTH CTAs once these MMAs have read it
            if (PROF) {
              const long long dt = NSK_CLK() - ci0;
              if (sh == SH_FP) t_iss_fp += dt; else if (sh == SH_Z || sh == SH_Z0) t_iss_z += dt; else t_iss_map += dt;
            }
            if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
          };
#pragma unroll 1
          for (int sg = 0; sg < nfull; ++sg) stage(std::integral_constant<int, (kps >> 4)>{}, sg == 0 ? 0u : 1u);
          if constexpr (ktail != 0) stage(std::integral_constant<int, (ktail >> 4)>{}, nfull == 0 ? 0u : 1u);
          umma_commit2(bars + 8 * op.commit0);
          if constexpr (op.commit1 != NOB) umma_commit2(bars + 8 * op.commit1);
        }
      };
      for (int64_t tp = tp0; tp < P.n_tp; tp += tp_step) for_each_op(issue_op, std::make_integer_sequence<int, NUM_OPS>{});
    }
    __syncwarp();
    if (PROF && P.prof && lane == 0) {
      P.prof[blockIdx.x * 16 + 2] = NSK_CLK() - t_begin; P.prof[blockIdx.x * 16 + 3] = t_dep; P.prof[blockIdx.x * 16 + 4] = t_w;
      P.prof[blockIdx.x * 16 + 5] = t_dep_map;
      P.prof[blockIdx.x * 16 + 13] = t_iss_map; P.prof[blockIdx.x * 16 + 14] = t_iss_fp; P.prof[blockIdx.x * 16 + 15] = t_iss_z;
    }
  } else if (warp >= EPI_WARP0 && warp < PRO_WARP0) {
    // ================================ epilogue ================================
    const int e = warp - EPI_WARP0;
    const int q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph_acca = 0, ph_mapb = 0, ph_fp0 = 0, ph_fp1 = 0;
    long long t_wmap = 0, t_wz = 0, t_wfp = 0, t_tail = 0;
    const long long t_begin = NSK_CLK();
    int par = 0;
    for (int64_t tp = tp0; tp < P.n_tp; tp += tp_step, par ^= 1) {
      const int64_t tile = tp * 2 + rank;
      // ---- mapping layers: LeakyReLU(acc) -> ACT_M (bias already inside the accumulator) ----
      for (int i = 1; i <= 5; ++i) {
        const bool fromB = (i & 1) == 0;
        const long long c0 = NSK_CLK();
        if (fromB) { mbar_wait(bars + 8 * B_MAPB, ph_mapb); ph_mapb ^= 1; }
        else { mbar_wait(bars + 8 * B_ACCA, ph_acca); ph_acca ^= 1; }
        t_wmap += NSK_CLK() - c0;
        tc_fence_after();
        const uint32_t src = tmem + (fromB ? TM_ACC_B : TM_ACC_A) + lane_off + hsel * 128;
        uint8_t* dst = smem + OFF_ACT_M + (uint32_t)(hsel * 16) * (TM * 16) + row * 16;
        uint32_t v[2][16];
        tmem_ld16(src, v[0]);
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < 8) tmem_ld16(src + (cb + 1) * 16, v[(cb + 1) & 1]);   // prefetch the next 16 columns
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float a0 = __uint_as_float(v[cb & 1][j]), a1 = __uint_as_float(v[cb & 1][j + 1]);
            pk[j >> 1] = pack_h2(fmaxf(a0, 0.2f * a0), fmaxf(a1, 0.2f * a1));   // LeakyReLU(0.2)
          }
          if (DBG && P.dbg && tile == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a0 = __uint_as_float(v[cb & 1][j]);
              P.dbg[((i - 1) * 128 + row) * 256 + hsel * 128 + cb * 16 + j] = fmaxf(a0, 0.2f * a0);
            }
          }
          *reinterpret_cast<uint4*>(dst + (cb * 2) * (TM * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dst + (cb * 2 + 1) * (TM * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        arrive_leader(bars + 8 * B_MACT);
      }
      // ---- trunk: h = sin(freq' * z + phase') ----
      float fin = 0.f;
      // one trunk layer; LAST (a compile-time tag) = the fifth layer, whose activations go into the 256 -> 1 dot product instead of
      // shared memory: peeled so that neither path carries the other's branch and code in its 8-column blocks
      auto trunk_layer = [&](auto last_tag, const int l) {
        constexpr bool LAST = decltype(last_tag)::value;
        long long c0 = NSK_CLK();
        mbar_wait(bars + 8 * B_ACCA, ph_acca); ph_acca ^= 1;   // Z_l
        t_wz += NSK_CLK() - c0;
        for (int c = 0; c < 4; ++c) {
          c0 = NSK_CLK();
          if (c & 1) { mbar_wait(bars + 8 * (B_FPFULL + 1), ph_fp1); ph_fp1 ^= 1; }
          else { mbar_wait(bars + 8 * (B_FPFULL + 0), ph_fp0); ph_fp0 ^= 1; }
          t_wfp += NSK_CLK() - c0;
          tc_fence_after();
          const uint32_t fp = tmem + ((c & 1) ? TM_FP1 : TM_FP0) + lane_off + hsel * 32;
          const uint32_t zz = tmem + TM_ACC_A + lane_off + c * 64 + hsel * 32;
          const int colw = c * 64 + hsel * 32;   // first layer column handled by this warp in this chunk
          uint32_t z[2][8], f[2][8], p[2][8];
          tmem_ld8(zz, z[0]); tmem_ld8(fp, f[0]); tmem_ld8(fp + 64, p[0]);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            tmem_ld_wait();
            if (pc + 1 < 4) {
              tmem_ld8(zz + (pc + 1) * 8, z[(pc + 1) & 1]);
              tmem_ld8(fp + (pc + 1) * 8, f[(pc + 1) & 1]);
              tmem_ld8(fp + 64 + (pc + 1) * 8, p[(pc + 1) & 1]);
            }
            float h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float arg = fmaf(__uint_as_float(f[pc & 1][j]), __uint_as_float(z[pc & 1][j]), __uint_as_float(p[pc & 1][j]));
              h[j] = (VARIANT & 1) ? arg : __sinf(arg);
            }
            if (DBG && P.dbg && tile == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) P.dbg[((5 + l) * 128 + row) * 256 + colw + pc * 8 + j] = h[j];
            }
            if (!LAST) {
              uint8_t* dst = smem + OFF_ACT_H + (uint32_t)((colw >> 3) + pc) * (TM * 16) + row * 16;
              *reinterpret_cast<uint4*>(dst) = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            } else {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(tailw + colw + pc * 8));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(tailw + colw + pc * 8 + 4));
              fin = fmaf(h[0], w0.x, fin); fin = fmaf(h[1], w0.y, fin); fin = fmaf(h[2], w0.z, fin); fin = fmaf(h[3], w0.w, fin);
              fin = fmaf(h[4], w1.x, fin); fin = fmaf(h[5], w1.y, fin); fin = fmaf(h[6], w1.z, fin); fin = fmaf(h[7], w1.w, fin);  // film_siren.py:147
            }
          }
          if (!LAST) fence_proxy_async_smem();
          tc_fence_before();
          arrive_leader(bars + 8 * (B_FPFREE + (c & 1)));
        }
      };
      for (int l = 0; l + 1 < DDF_LAYERS; ++l) trunk_layer(std::false_type{}, l);
      trunk_layer(std::true_type{}, DDF_LAYERS - 1);
      // ---- tail: sigmoid, visibility, Lambertian accumulation ----
      const long long c_tail = NSK_CLK();
      if (hsel == 1) fin_part[row] = fin;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (hsel == 0) {
        const int64_t pr = tile * TM + row;
        const bool valid = pr < P.n_pairs;
        const int64_t prc = valid ? pr : P.n_pairs - 1;
        const int64_t ray = prc / P.Dp;
        const int j = (int)(prc % P.Dp);
        const float o = fin + fin_part[row] + __ldg(tailw + 256);
        const float ddf = sigmoidf_(o) * (2.0f * P.radius);       // directional_distance_field.py:297-299
        const float term = geo_term[par * 128 + row];
        const float vis = visibility_from_ddf(ddf, term, P.radius, P.thr, P.sig_scale);
        if (valid) {
          if (P.vis_out) P.vis_out[pr] = vis;
          if (P.ddf_out) P.ddf_out[pr] = ddf;
          if (P.term_out) P.term_out[pr] = term;
        }
        const float lx = __ldg(P.dirs + j * 3), ly = __ldg(P.dirs + j * 3 + 1), lz = __ldg(P.dirs + j * 3 + 2);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (P.S == 0) {
          // pre-collapsed shading coefficients G [R, Dp, 3] = sum_s wa * clamp01(n.l_j) * inv_count (nsk_lambert_collapse_sel):
          // full renders have 48..128 samples per ray, and running that sum here costs the tile ~20 % (it is per pair, on 4 warps)
          const float* gp = P.wa + prc * 3;
          c0 = __ldg(gp); c1 = __ldg(gp + 1); c2 = __ldg(gp + 2);
        }
        for (int s = 0; s < P.S; ++s) {
          const int64_t i = ray * P.S + s;
          float c = __ldg(P.normals + i * 3) * lx + __ldg(P.normals + i * 3 + 1) * ly + __ldg(P.normals + i * 3 + 2) * lz;
          c = fminf(fmaxf(c, 0.f), 1.f) * __ldg(P.inv_count + i);
          c0 = fmaf(__ldg(P.wa + i * 3), c, c0); c1 = fmaf(__ldg(P.wa + i * 3 + 1), c, c1); c2 = fmaf(__ldg(P.wa + i * 3 + 2), c, c2);
        }
        const float* rad = P.radiance + ((int64_t)(P.cam ? P.cam[ray] : 0) * P.Dp + j) * 3;
        const float k = valid ? vis : 0.f;
        c0 *= k * __ldg(rad); c1 *= k * __ldg(rad + 1); c2 *= k * __ldg(rad + 2);
        // rows of a tile mostly share one ray: reduce across the warp when they do
        const int64_t ray0 = __shfl_sync(0xffffffffu, ray, 0);
        if (__all_sync(0xffffffffu, ray == ray0)) {
          c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
          if (lane == 0) { atomicAdd(P.rgb_lin + ray * 3, c0); atomicAdd(P.rgb_lin + ray * 3 + 1, c1); atomicAdd(P.rgb_lin + ray * 3 + 2, c2); }
        } else if (valid) {
          atomicAdd(P.rgb_lin + ray * 3, c0); atomicAdd(P.rgb_lin + ray * 3 + 1, c1); atomicAdd(P.rgb_lin + ray * 3 + 2, c2);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // fin_part reusable
      t_tail += NSK_CLK() - c_tail;
    }
    if (P.prof && e == 0 && lane == 0) {
      P.prof[blockIdx.x * 16 + 6] = NSK_CLK() - t_begin; P.prof[blockIdx.x * 16 + 7] = t_wmap; P.prof[blockIdx.x * 16 + 8] = t_wz;
      P.prof[blockIdx.x * 16 + 9] = t_wfp; P.prof[blockIdx.x * 16 + 10] = t_tail;
    }
  } else if (warp >= PRO_WARP0) {
    // ================================ prologue ================================
    const int row = (warp - PRO_WARP0) * 32 + lane;
    const uint32_t mask = (1u << P.log2_T) - 1u;
    uint32_t ph_empty = 0;
    int par = 0;
    long long t_wempty = 0;
    const long long t_begin = NSK_CLK();
    for (int64_t tp = tp0; tp < P.n_tp; tp += tp_step, par ^= 1) {
      const int64_t tile = tp * 2 + rank;
      const int64_t pr = min(tile * TM + row, P.n_pairs - 1);
      const int64_t ray = pr / P.Dp;
      const int j = (int)(pr % P.Dp);
      const float p[3] = {__ldg(P.points + ray * 3), __ldg(P.points + ray * 3 + 1), __ldg(P.points + ray * 3 + 2)};
      const float l[3] = {__ldg(P.dirs + j * 3), __ldg(P.dirs + j * 3 + 1), __ldg(P.dirs + j * 3 + 2)};
      float qv[3], tt;
      sphere_exit(p, l, P.radius, qv, tt);                                   // neusky_model.py:1693
      const float dx = qv[0] - p[0], dy = qv[1] - p[1], dz = qv[2] - p[2];
      const float term = sqrtf(dx * dx + dy * dy + dz * dz);                 // neusky_model.py:1697
      const float dneg[3] = {-l[0], -l[1], -l[2]};                            // neusky_model.py:1702
      float dl[3], feat[16];
      ddf_local_dir(qv, dneg, dl);                                           // ddf_model.py:158-200
      ddf_dir_features(dl, feat);                                            // directional_distance_field.py:270-271
      feat[15] = 0.f;
      // mapping input: [q (3) | hash(q) (32) | 1, 1 (bias columns) | zero pad] = 48 halves
      uint32_t mp[24];
      float carry = qv[2];   // element 2 pairs with the first hash feature
      mp[0] = pack_h2(qv[0], qv[1]);
#pragma unroll
      for (int lev = 0; lev < DDF_LEVELS; lev += 2) {
        float2 f[2][8];
        float ox[2], oy[2], oz[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t idx[8];
          float dw_[3], sc_;
          grid_corners(P.gm, lev + u, qv[0], qv[1], qv[2], __ldg(P.scalings + lev + u), mask, idx, ox[u], oy[u], oz[u], dw_, sc_);
          const float2* tl = P.table + ((size_t)(lev + u) << P.log2_T);
#pragma unroll
          for (int c = 0; c < 8; ++c) f[u][c] = __ldg(tl + idx[c]);
        }
        const float2 r0 = hash_interp(f[0], ox[0], oy[0], oz[0]);
        const float2 r1 = hash_interp(f[1], ox[1], oy[1], oz[1]);
        // elements 3+2*lev .. 3+2*lev+3 ; packed pairs start at odd element indices
        mp[1 + lev] = pack_h2(carry, r0.x);
        mp[2 + lev] = pack_h2(r0.y, r1.x);
        carry = r1.y;
      }
      mp[17] = pack_h2(carry, 1.0f);     // elements 34, 35 (35 = bias hi column)
      mp[18] = pack_h2(1.0f, 0.0f);      // element 36 = bias lo column
      mp[19] = mp[20] = mp[21] = mp[22] = mp[23] = 0u;
      // wait until the MMAs of the previous tile have consumed IN_M / IN_H
      {
        const long long c0 = NSK_CLK();
        mbar_wait(bars + 8 * B_INEMPTY, ph_empty ^ 1); ph_empty ^= 1;
        t_wempty += NSK_CLK() - c0;
      }
      {
        uint8_t* dm = smem + OFF_IN_M + row * 16;
#pragma unroll
        for (int kc = 0; kc < 6; ++kc)
          *reinterpret_cast<uint4*>(dm + kc * (TM * 16)) = make_uint4(mp[kc * 4], mp[kc * 4 + 1], mp[kc * 4 + 2], mp[kc * 4 + 3]);
        // trunk input [x_hi (16) | x_lo (16) | x_hi (16)]: x = hi + lo exactly to ~2^-22, both halves fp16
        uint8_t* dh = smem + OFF_IN_H + row * 16;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = feat[kc * 8 + 2 * j], b = feat[kc * 8 + 2 * j + 1];
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j] = pack_h2(a - hf.x, b - hf.y);
          }
          const uint4 vh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dh + kc * (TM * 16)) = vh;
          *reinterpret_cast<uint4*>(dh + (2 + kc) * (TM * 16)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(dh + (4 + kc) * (TM * 16)) = vh;
        }
      }
      geo_term[par * 128 + row] = term;
      fence_proxy_async_smem();
      arrive_leader(bars + 8 * B_INFULL);
    }
    if (P.prof && row == 0) { P.prof[blockIdx.x * 16 + 11] = NSK_CLK() - t_begin; P.prof[blockIdx.x * 16 + 12] = t_wempty; }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA may exit (or free TMEM) while the other can still touch its barriers / operands
  if (warp == 2) tmem_dealloc2<512>(tmem);
}

}  // namespace tcs2
}  // namespace nsk

extern "C" int64_t nsk_ddf_tc2_weights_bytes(void) { return nsk::tcs2::BLOB_BYTES; }

// Diagnostics only (not part of the public ABI): when set, the first tile of CTA 0 dumps its
// post-activation values [10][128][256] (5 mapping layers, 5 trunk layers) into this device buffer.
static float* g_nsk_tc2_debug_dump = nullptr;
extern "C" void nsk_debug_set_tc2_dump(float* buf) { g_nsk_tc2_debug_dump = buf; }
// Diagnostics only: per-CTA cycle counters [grid][16] (see scripts/k4_phase_profile.py for the slot meaning).
static unsigned long long* g_nsk_tc2_prof = nullptr;
extern "C" void nsk_debug_set_tc2_prof(unsigned long long* buf) { g_nsk_tc2_prof = buf; }

extern "C" int nsk_sky_shade_tc2_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                                    const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                                    const int32_t* cam, const void* ddf_weights, const float* hash_table,
                                    const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                                    float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out, float* term_out,
                                    void* stream) {
  return nsk_sky_shade_tc2_fwd_ex(points, R, normals, wa, inv_count, S, dirs, Dp, radiance, cam, ddf_weights, hash_table, scalings, num_levels, log2_T,
                                  nullptr, 0, radius, threshold, sigmoid_scale, rgb_lin, vis_out, ddf_out, term_out, stream);
}

extern "C" int nsk_sky_shade_tc2_fwd_ex(const float* points, int64_t R, const float* normals, const float* wa,
                                       const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                                       const int32_t* cam, const void* ddf_weights, const float* hash_table,
                                       const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                                       float radius, float threshold, float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out,
                                       float* term_out, void* stream) {
  using namespace nsk::tcs2;
  NSK_REQUIRE(grid_meta == nullptr || (reinterpret_cast<uintptr_t>(grid_meta) & 15) == 0, "nsk_sky_shade_tc2_fwd_ex: grid_meta must be 16-byte aligned");
  NSK_REQUIRE(num_levels == nsk::DDF_LEVELS, "nsk_sky_shade_tc2_fwd: the DDF position encoding has 16 levels");
  if (R == 0 || Dp == 0) return 0;
  NSK_REQUIRE(S >= 0, "nsk_sky_shade_tc2_fwd: S must be >= 0");
  NSK_REQUIRE(points && wa && dirs && radiance && ddf_weights && hash_table && scalings && rgb_lin && (S == 0 || (normals && inv_count)),
              "nsk_sky_shade_tc2_fwd: null pointer");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(ddf_weights) & 15) == 0, "nsk_sky_shade_tc2_fwd: weight blob must be 16-byte aligned");
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "nsk_sky_shade_tc2_fwd: device setup", &num_sms, [] {
        cudaError_t e = cudaFuncSetAttribute(sky_shade_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sky_shade_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sky_shade_tc2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        return e;
      }))
    return err;
  Params P;
  P.points = points; P.R = R; P.normals = normals; P.wa = wa; P.inv_count = inv_count; P.S = S;
  P.dirs = dirs; P.Dp = Dp; P.radiance = radiance; P.cam = cam;
  P.blob = reinterpret_cast<const uint8_t*>(ddf_weights);
  P.table = reinterpret_cast<const float2*>(hash_table); P.scalings = scalings; P.log2_T = log2_T;
  P.radius = radius; P.thr = threshold; P.sig_scale = sigmoid_scale;
  P.rgb_lin = rgb_lin; P.vis_out = vis_out; P.ddf_out = ddf_out; P.term_out = term_out;
  P.gm = nsk::GridMode{reinterpret_cast<const int4*>(grid_meta), smoothstep};
  P.dbg = g_nsk_tc2_debug_dump;
  P.prof = g_nsk_tc2_prof;
  P.n_pairs = R * (int64_t)Dp;
  P.n_tiles = (P.n_pairs + TM - 1) / TM;
  P.n_tp = (P.n_tiles + 1) / 2;
  const int64_t clusters = P.n_tp < num_sms / 2 ? P.n_tp : num_sms / 2;
  const int64_t grid = 2 * clusters;
  if (P.prof) sky_shade_tc2_kernel<4><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  else if (P.dbg) sky_shade_tc2_kernel<8><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  else sky_shade_tc2_kernel<0><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  return nsk::check_launch("sky_shade_tc2_kernel");
}

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
//   SMEM   ACT_M 64 KB  mapping activations / m5 (A operand, fp16, K-major no-swizzle canonical layout)
//          ACT_H 64 KB  trunk activations h_l
//          IN_M 16 KB, IN_H 8 KB  first-layer inputs of the NEXT tile (written by the prologue warps)
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

namespace nsk {
namespace tcs {

using namespace nsk::tc;

constexpr int TM = 128;                      // rows (pairs) per tile
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4, PRO_WARP0 = 12;
constexpr int EPI_THREADS = 256, PRO_THREADS = 128;

// stream: M1 (2 stages) | M2..M5 (8 each) | per layer: FP(l,0) 4, Z_l (1 or 8), FP(l,1..3) 4 each
constexpr int STAGES_PER_TILE = 2 + 4 * 8 + (1 + 4 * 8) + 20 * 4;  // 147
constexpr int BIAS_FLOATS = 5 * 256 /*map*/ + 5 * 256 /*freq'*/ + 5 * 256 /*phase'*/ + 256 /*w_final*/ + 4;
constexpr int64_t BLOB_BYTES = (int64_t)STAGES_PER_TILE * STAGE_BYTES + (int64_t)BIAS_FLOATS * 4;

// shared memory carve-up (bytes)
constexpr uint32_t OFF_ACT_M = 0;
constexpr uint32_t OFF_ACT_H = 65536;
constexpr uint32_t OFF_IN_M = 131072;              // [128][64] fp16
constexpr uint32_t OFF_IN_H = OFF_IN_M + 16384;    // [128][32] fp16
constexpr uint32_t OFF_RING = OFF_IN_H + 8192;
constexpr uint32_t OFF_GEO = OFF_RING + NSTAGE * STAGE_BYTES;   // [2][128] x {term, pad} + fin[128]
constexpr uint32_t OFF_BAR = OFF_GEO + 2 * 128 * 4 + 128 * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32 * 8 + 16;

// barrier indices
enum { B_WFULL = 0, B_WEMPTY = 4, B_INFULL = 8, B_INEMPTY = 9, B_ACCA = 10, B_MAPB = 11, B_FPFULL = 12, B_MACT = 14, B_FPFREE = 15, B_COUNT = 17 };

constexpr uint32_t TM_ACC_A = 0, TM_ACC_B = 256, TM_FP0 = 256, TM_FP1 = 384;

struct Params {
  const float* points; int64_t R;
  const float* normals; const float* wa; const float* inv_count; int S;
  const float* dirs; int Dp;
  const float* radiance; const int32_t* cam;
  const uint8_t* blob; const float2* table; const float* scalings; int log2_T;
  float radius, thr, sig_scale;
  float* rgb_lin; float* vis_out; float* ddf_out; float* term_out;
  int64_t n_pairs, n_tiles;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Issue the MMAs of one GEMM op: D[tmem_d, 128 x N] (+)= A[smem a_base, 128 x K] * W^T, with W arriving
// in `nst` ring stages of `kps` K-columns each ([N][kps] canonical tiles).
__device__ __forceinline__ void issue_op(uint32_t smem_base, uint32_t bars, uint32_t a_off, int N, int nst, int kps,
                                         uint32_t tmem_d, uint32_t& wstage, uint32_t& wphase) {
  const uint32_t idesc = make_idesc_f16(TM, N);
  uint32_t acc = 0;
  for (int s = 0; s < nst; ++s) {
    mbar_wait(bars + 8 * (B_WFULL + wstage), wphase);
    tc_fence_after();
    const uint32_t b_base = smem_base + OFF_RING + wstage * STAGE_BYTES;
    for (int j = 0; j < kps / 16; ++j) {
      const int k0 = s * kps + j * 16;
      const uint64_t ad = make_smem_desc(smem_base + a_off + (uint32_t)(k0 / 8) * (TM * 16), TM * 16, 128);
      const uint64_t bd = make_smem_desc(b_base + (uint32_t)(j * 2) * (uint32_t)(N * 16), (uint32_t)(N * 16), 128);
      umma_ss(tmem_d, ad, bd, idesc, acc);
      acc = 1;
    }
    umma_commit(bars + 8 * (B_WEMPTY + wstage));   // stage free once these MMAs have read it
    if (++wstage == NSTAGE) { wstage = 0; wphase ^= 1; }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) sky_shade_tc_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32 * 8);
  float* geo_term = reinterpret_cast<float*>(smem + OFF_GEO);           // [2][128]
  float* fin_part = reinterpret_cast<float*>(smem + OFF_GEO + 2 * 128 * 4);  // [128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* bias = reinterpret_cast<const float*>(P.blob + (size_t)STAGES_PER_TILE * STAGE_BYTES);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bars + 8 * (B_WFULL + i), 1); mbar_init(bars + 8 * (B_WEMPTY + i), 1); }
    mbar_init(bars + 8 * B_INFULL, PRO_THREADS);
    mbar_init(bars + 8 * B_INEMPTY, 1);
    mbar_init(bars + 8 * B_ACCA, 1);
    mbar_init(bars + 8 * B_MAPB, 1);
    mbar_init(bars + 8 * (B_FPFULL + 0), 1); mbar_init(bars + 8 * (B_FPFULL + 1), 1);
    mbar_init(bars + 8 * B_MACT, EPI_THREADS);
    mbar_init(bars + 8 * (B_FPFREE + 0), EPI_THREADS); mbar_init(bars + 8 * (B_FPFREE + 1), EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_last();
      uint32_t st = 0, ph = 0;
      for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (int i = 0; i < STAGES_PER_TILE; ++i) {
          mbar_wait(bars + 8 * (B_WEMPTY + st), ph ^ 1);
          mbar_arrive_expect_tx(bars + 8 * (B_WFULL + st), STAGE_BYTES);
          bulk_g2s_hint(sbase + OFF_RING + st * STAGE_BYTES, P.blob + (size_t)i * STAGE_BYTES, STAGE_BYTES, bars + 8 * (B_WFULL + st), pol);
          if (++st == NSTAGE) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      uint32_t wst = 0, wph = 0;
      uint32_t ph_in = 0, ph_mact = 0, ph_free0 = 0, ph_free1 = 0;
      for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        mbar_wait(bars + 8 * B_INFULL, ph_in); ph_in ^= 1;
        tc_fence_after();
        // ---- mapping network ----
        issue_op(sbase, bars, OFF_IN_M, 256, 2, 32, tmem + TM_ACC_A, wst, wph);
        umma_commit(bars + 8 * B_ACCA);
        for (int i = 2; i <= 5; ++i) {
          mbar_wait(bars + 8 * B_MACT, ph_mact); ph_mact ^= 1;
          tc_fence_after();
          const bool toB = (i & 1) == 0;
          issue_op(sbase, bars, OFF_ACT_M, 256, 8, 32, tmem + (toB ? TM_ACC_B : TM_ACC_A), wst, wph);
          umma_commit(bars + 8 * (toB ? B_MAPB : B_ACCA));
        }
        mbar_wait(bars + 8 * B_MACT, ph_mact); ph_mact ^= 1;   // m5 ready, ACC_A / ACC_B free
        tc_fence_after();
        // ---- trunk ----
        issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP0, wst, wph);       // FP(0,0)
        umma_commit(bars + 8 * (B_FPFULL + 0));
        issue_op(sbase, bars, OFF_IN_H, 256, 1, 32, tmem + TM_ACC_A, wst, wph);      // Z_0
        umma_commit(bars + 8 * B_ACCA);
        umma_commit(bars + 8 * B_INEMPTY);                                           // IN_M / IN_H consumed
        issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP1, wst, wph);       // FP(0,1)
        umma_commit(bars + 8 * (B_FPFULL + 1));
        for (int l = 0; l < DDF_LAYERS; ++l) {
          mbar_wait(bars + 8 * (B_FPFREE + 0), ph_free0); ph_free0 ^= 1;             // C(l,0) done
          tc_fence_after();
          issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP0, wst, wph);     // FP(l,2)
          umma_commit(bars + 8 * (B_FPFULL + 0));
          mbar_wait(bars + 8 * (B_FPFREE + 1), ph_free1); ph_free1 ^= 1;             // C(l,1) done
          tc_fence_after();
          issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP1, wst, wph);     // FP(l,3)
          umma_commit(bars + 8 * (B_FPFULL + 1));
          mbar_wait(bars + 8 * (B_FPFREE + 0), ph_free0); ph_free0 ^= 1;             // C(l,2) done
          tc_fence_after();
          if (l + 1 < DDF_LAYERS) {
            issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP0, wst, wph);   // FP(l+1,0)
            umma_commit(bars + 8 * (B_FPFULL + 0));
          }
          mbar_wait(bars + 8 * (B_FPFREE + 1), ph_free1); ph_free1 ^= 1;             // C(l,3) done: h_l complete
          tc_fence_after();
          if (l + 1 < DDF_LAYERS) {
            issue_op(sbase, bars, OFF_ACT_H, 256, 8, 32, tmem + TM_ACC_A, wst, wph); // Z_{l+1}
            umma_commit(bars + 8 * B_ACCA);
            issue_op(sbase, bars, OFF_ACT_M, 128, 4, 64, tmem + TM_FP1, wst, wph);   // FP(l+1,1)
            umma_commit(bars + 8 * (B_FPFULL + 1));
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < PRO_WARP0) {
    // ================================ epilogue ================================
    const int e = warp - EPI_WARP0;
    const int q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph_acca = 0, ph_mapb = 0, ph_fp0 = 0, ph_fp1 = 0;
    int par = 0;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, par ^= 1) {
      // ---- mapping layers: LeakyReLU(acc + b) -> ACT_M ----
      for (int i = 1; i <= 5; ++i) {
        const bool fromB = (i & 1) == 0;
        if (fromB) { mbar_wait(bars + 8 * B_MAPB, ph_mapb); ph_mapb ^= 1; }
        else { mbar_wait(bars + 8 * B_ACCA, ph_acca); ph_acca ^= 1; }
        tc_fence_after();
        const float* b = bias + (i - 1) * 256;
        const uint32_t src = tmem + (fromB ? TM_ACC_B : TM_ACC_A) + lane_off + hsel * 128;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
          uint32_t v[16];
          tmem_ld16(src + cb * 16, v);
          tmem_ld_wait();
          const int col0 = hsel * 128 + cb * 16;
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float a0 = __uint_as_float(v[j]) + __ldg(b + col0 + j);
            float a1 = __uint_as_float(v[j + 1]) + __ldg(b + col0 + j + 1);
            a0 = a0 > 0.f ? a0 : 0.2f * a0;
            a1 = a1 > 0.f ? a1 : 0.2f * a1;
            pk[j >> 1] = pack_h2(a0, a1);
          }
          uint8_t* dst = smem + OFF_ACT_M + (uint32_t)(col0 >> 3) * (TM * 16) + row * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dst + TM * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bars + 8 * B_MACT);
      }
      // ---- trunk: h = sin(freq' * z + phase') ----
      float fin = 0.f;
      for (int l = 0; l < DDF_LAYERS; ++l) {
        mbar_wait(bars + 8 * B_ACCA, ph_acca); ph_acca ^= 1;   // Z_l
        const float* bf = bias + 5 * 256 + l * 256;
        const float* bp = bias + 10 * 256 + l * 256;
        for (int c = 0; c < 4; ++c) {
          if (c & 1) { mbar_wait(bars + 8 * (B_FPFULL + 1), ph_fp1); ph_fp1 ^= 1; }
          else { mbar_wait(bars + 8 * (B_FPFULL + 0), ph_fp0); ph_fp0 ^= 1; }
          tc_fence_after();
          const uint32_t fp = tmem + ((c & 1) ? TM_FP1 : TM_FP0) + lane_off;
#pragma unroll 1
          for (int sb = 0; sb < 2; ++sb) {
            const int cc = hsel * 32 + sb * 16;          // column inside the 64-wide chunk
            const int col0 = c * 64 + cc;                // column inside the layer
            uint32_t z[16], f[16], p[16];
            tmem_ld16(tmem + TM_ACC_A + lane_off + col0, z);
            tmem_ld16(fp + cc, f);
            tmem_ld16(fp + 64 + cc, p);
            tmem_ld_wait();
            float h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float fr = __uint_as_float(f[j]) + __ldg(bf + col0 + j);
              const float ph = __uint_as_float(p[j]) + __ldg(bp + col0 + j);
              h[j] = __sinf(fmaf(fr, __uint_as_float(z[j]), ph));
            }
            if (l + 1 < DDF_LAYERS) {
              uint8_t* dst = smem + OFF_ACT_H + (uint32_t)(col0 >> 3) * (TM * 16) + row * 16;
              *reinterpret_cast<uint4*>(dst) = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
              *reinterpret_cast<uint4*>(dst + TM * 16) = make_uint4(pack_h2(h[8], h[9]), pack_h2(h[10], h[11]), pack_h2(h[12], h[13]), pack_h2(h[14], h[15]));
            } else {
              const float* wf = bias + 15 * 256;
#pragma unroll
              for (int j = 0; j < 16; ++j) fin = fmaf(h[j], __ldg(wf + col0 + j), fin);   // final 256 -> 1 (film_siren.py:147)
            }
          }
          if (l + 1 < DDF_LAYERS) fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(bars + 8 * (B_FPFREE + (c & 1)));
        }
      }
      // ---- tail: sigmoid, visibility, Lambertian accumulation ----
      if (hsel == 1) fin_part[row] = fin;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (hsel == 0) {
        const int64_t pr = tile * TM + row;
        const bool valid = pr < P.n_pairs;
        const int64_t prc = valid ? pr : P.n_pairs - 1;
        const int64_t ray = prc / P.Dp;
        const int j = (int)(prc % P.Dp);
        const float o = fin + fin_part[row] + __ldg(bias + 16 * 256);
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
    }
  } else if (warp >= PRO_WARP0) {
    // ================================ prologue ================================
    const int row = (warp - PRO_WARP0) * 32 + lane;
    const uint32_t mask = (1u << P.log2_T) - 1u;
    uint32_t ph_empty = 0;
    int par = 0;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, par ^= 1) {
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
      // mapping input: [q (3) | hash(q) (32) | zero pad] = 64 halves
      float mi[40];
      mi[0] = qv[0]; mi[1] = qv[1]; mi[2] = qv[2];
#pragma unroll 1
      for (int lev = 0; lev < DDF_LEVELS; lev += 2) {
        float2 f[2][8];
        float ox[2], oy[2], oz[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float s = __ldg(P.scalings + lev + u);
          uint32_t idx[8];
          hash_corners(__fmul_rn(qv[0], s), __fmul_rn(qv[1], s), __fmul_rn(qv[2], s), mask, idx, ox[u], oy[u], oz[u]);
          const float2* tl = P.table + ((size_t)(lev + u) << P.log2_T);
#pragma unroll
          for (int c = 0; c < 8; ++c) f[u][c] = __ldg(tl + idx[c]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float2 r = hash_interp(f[u], ox[u], oy[u], oz[u]);
          mi[3 + 2 * (lev + u)] = r.x;
          mi[3 + 2 * (lev + u) + 1] = r.y;
        }
      }
      mi[35] = mi[36] = mi[37] = mi[38] = mi[39] = 0.f;
      // wait until the MMAs of the previous tile have consumed IN_M / IN_H
      mbar_wait(bars + 8 * B_INEMPTY, ph_empty ^ 1); ph_empty ^= 1;
      {
        uint8_t* dm = smem + OFF_IN_M + row * 16;
#pragma unroll
        for (int kc = 0; kc < 5; ++kc)
          *reinterpret_cast<uint4*>(dm + kc * (TM * 16)) = make_uint4(pack_h2(mi[kc * 8], mi[kc * 8 + 1]), pack_h2(mi[kc * 8 + 2], mi[kc * 8 + 3]),
                                                                       pack_h2(mi[kc * 8 + 4], mi[kc * 8 + 5]), pack_h2(mi[kc * 8 + 6], mi[kc * 8 + 7]));
#pragma unroll
        for (int kc = 5; kc < 8; ++kc) *reinterpret_cast<uint4*>(dm + kc * (TM * 16)) = make_uint4(0, 0, 0, 0);
        uint8_t* dh = smem + OFF_IN_H + row * 16;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
          *reinterpret_cast<uint4*>(dh + kc * (TM * 16)) = make_uint4(pack_h2(feat[kc * 8], feat[kc * 8 + 1]), pack_h2(feat[kc * 8 + 2], feat[kc * 8 + 3]),
                                                                       pack_h2(feat[kc * 8 + 4], feat[kc * 8 + 5]), pack_h2(feat[kc * 8 + 6], feat[kc * 8 + 7]));
#pragma unroll
        for (int kc = 2; kc < 4; ++kc) *reinterpret_cast<uint4*>(dh + kc * (TM * 16)) = make_uint4(0, 0, 0, 0);
      }
      geo_term[par * 128 + row] = term;
      fence_proxy_async_smem();
      mbar_arrive(bars + 8 * B_INFULL);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace tcs
}  // namespace nsk

extern "C" int64_t nsk_ddf_tc_weights_bytes(void) { return nsk::tcs::BLOB_BYTES; }

extern "C" int nsk_sky_shade_tc_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                                    const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                                    const int32_t* cam, const void* ddf_weights, const float* hash_table,
                                    const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                                    float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out, float* term_out,
                                    void* stream) {
  using namespace nsk::tcs;
  NSK_REQUIRE(num_levels == nsk::DDF_LEVELS, "nsk_sky_shade_tc_fwd: the DDF position encoding has 16 levels");
  if (R == 0 || Dp == 0) return 0;
  NSK_REQUIRE(S >= 1, "nsk_sky_shade_tc_fwd: S must be >= 1");
  NSK_REQUIRE(points && normals && wa && inv_count && dirs && radiance && ddf_weights && hash_table && scalings && rgb_lin,
              "nsk_sky_shade_tc_fwd: null pointer");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(ddf_weights) & 15) == 0, "nsk_sky_shade_tc_fwd: weight blob must be 16-byte aligned");
  static thread_local int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sky_shade_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) { num_sms = 0; return nsk::fail("nsk_sky_shade_tc_fwd: device setup", cudaGetErrorString(e)); }
  }
  Params P;
  P.points = points; P.R = R; P.normals = normals; P.wa = wa; P.inv_count = inv_count; P.S = S;
  P.dirs = dirs; P.Dp = Dp; P.radiance = radiance; P.cam = cam;
  P.blob = reinterpret_cast<const uint8_t*>(ddf_weights);
  P.table = reinterpret_cast<const float2*>(hash_table); P.scalings = scalings; P.log2_T = log2_T;
  P.radius = radius; P.thr = threshold; P.sig_scale = sigmoid_scale;
  P.rgb_lin = rgb_lin; P.vis_out = vis_out; P.ddf_out = ddf_out; P.term_out = term_out;
  P.n_pairs = R * (int64_t)Dp;
  P.n_tiles = (P.n_pairs + TM - 1) / TM;
  const int64_t grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  sky_shade_tc_kernel<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  return nsk::check_launch("sky_shade_tc_kernel");
}

// K2 (tensor-core path): SDF / albedo field of a ray sample on tcgen05 / TMEM, fused with its hash-grid encode and
// the ANALYTIC gradient d sdf / d x.
//
// Replaces SDFAlbedoField.get_outputs (neusky/fields/sdf_albedo_field.py:211-269): forward_geonetwork [NS-mem,
// SURVEY A.4] (:233: L-inf contraction, (p+2)/4, hash grid, [x | PE6(x) | feat] -> 256 -> 256 -> 1+256, softplus beta=100),
// torch.autograd.grad(sdf, x, create_graph) (:235-238) and get_colors (:185-209: [x | PE6(x) | geo] -> 256 -> 256 -> 3).
// The reference differentiates through a retained autograd graph; here the reverse pass is two more GEMMs
// (W1^T, W0^T) over activations that are still on chip, plus the closed-form derivative of the PE and of the trilinear
// hash interpolation.
//
// Numerics: fp16 operands, fp32 accumulation in TMEM, fp32 epilogues.  x enters layer 0 as fp16 hi + lo columns, the sdf
// output (last-layer row 0) is an fp32 dot product in the epilogue.  Tolerances for this path: tests/test_gpu_sdf.py.
//
// One persistent CTA per SM, tiles of 128 samples (rows).  Per tile, in MMA issue order (accumulators ACC_A = TMEM cols
// [0,256), ACC_B = [256,512); SMEM A operands IN [128][80], X [128][272], Y [128][272], fp16 K-major no-swizzle):
//   G0  IN x W0        -> A   E1: a0 = softplus(z0)                         -> X
//   G1  X  x W1        -> B   E2: a1 = softplus(z1) -> X ; g2 = W2[0,:]*sigmoid(100 z1) -> Y ; sdf = a1 . W2[0,:] + b (fp32)
//   G2  X  x W2[1:]    -> A   E3: geo feature -> X
//   B1  Y  x W1^T      -> B
//   G0' IN x W0        -> A   (z0 again: cheaper than keeping sigmoid(100 z0) for 128 rows on chip)
//                             E4: g1 = (W1^T g2) * sigmoid(100 z0)              -> Y
//   B0  Y  x W0^T      -> B[0,80)   read by the prologue warps: d sdf / d(x, PE, feat) -> chain rule -> grad [n,3]
//   C0  [IN(48) | X] x Wc0 -> A   E5: relu -> Y
//   C1  Y  x Wc1       -> B   E6: relu -> X
//   C2  X  x Wc2       -> A[0,16)  E7: sigmoid -> albedo [n,3]
// Every bias rides inside its GEMM as two extra K columns (fp16 hi + lo) against constant-1 activation columns.
//
//   warp 0      weight producer: cp.async.bulk of the pre-tiled fp16 weight stream (L2 resident) into a 4 x 16 KB ring
//   warp 1      MMA issuer (one elected lane, __constant__ segment table, mbarrier-gated)
//   warp 2      TMEM allocator
//   warps 4-19  epilogue (FOUR warps per TMEM lane quadrant = per scheduler, 64 columns per thread).  This chain is serial (every GEMM
//               waits for the epilogue before it), and its softplus / sigmoid epilogues are MUFU work: with two warps per scheduler the
//               XU delivers 9.1 results per clock and SM, with four 15.8 (tools/tc_probe.cu `sin`, profiles/r02_tmem_read_and_sin_probes.log)
//   warps 20-23 prologue (thread = row): contraction, 16-level hash gather, PE of the NEXT tile; gradient chain of the
//               current tile (reads d sdf / d input straight from TMEM, re-gathers the hash corners)
//   Registers: 768 threads x 80 at launch; setmaxnreg moves 32 per thread from the service warpgroup (48) to the prologue warpgroup (112)
#include "nsk_common.cuh"
#include "tc_util.cuh"
#include "sdf_common.cuh"

namespace nsk {
namespace sdftc {

using namespace nsk::tc;

constexpr int TM = 128;
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr int NUM_THREADS = 768;
constexpr int EPI_WARP0 = 4, PRO_WARP0 = 20;
constexpr int EPI_THREADS = 512, PRO_THREADS = 128;
constexpr int EPI_GROUPS = EPI_THREADS / 128;     // epilogue warps per TMEM lane quadrant (= per scheduler) = column groups of a row
constexpr int EPI_COLS = 256 / EPI_GROUPS;        // columns per epilogue thread
constexpr int KIN = 80, KX = 272;

// ---- weight stream (bytes), one region per GEMM; G0' re-reads the G0 region ------------------------------
constexpr int64_t SZ_G0 = 256 * 80 * 2, SZ_272 = 256 * 272 * 2, SZ_B1 = 256 * 256 * 2, SZ_B0 = 80 * 256 * 2, SZ_C0 = 256 * (48 + 272) * 2,
                  SZ_C2 = 16 * 272 * 2;
constexpr int64_t O_G0 = 0, O_G1 = O_G0 + SZ_G0, O_G2 = O_G1 + SZ_272, O_B1 = O_G2 + SZ_272, O_B0 = O_B1 + SZ_B1, O_C0 = O_B0 + SZ_B0,
                  O_C1 = O_C0 + SZ_C0, O_C2 = O_C1 + SZ_272, STREAM_BYTES = O_C2 + SZ_C2;
constexpr int TAIL_FLOATS = 256 /* W2[0,:] */ + 4 /* b2[0] */;
constexpr int64_t BLOB_BYTES = STREAM_BYTES + (int64_t)TAIL_FLOATS * 4;

// ---- shared memory ------------------------------------------------------------------------------------
constexpr uint32_t OFF_IN = 0;                               // [128][80]  fp16
constexpr uint32_t OFF_X = OFF_IN + TM * KIN * 2;            // [128][272] fp16
constexpr uint32_t OFF_Y = OFF_X + TM * KX * 2;              // [128][272] fp16
constexpr uint32_t OFF_RING = OFF_Y + TM * KX * 2;
constexpr uint32_t OFF_MISC = OFF_RING + NSTAGE * STAGE_BYTES;   // sdf partial sums [EPI_GROUPS - 1][128] fp32
constexpr uint32_t OFF_BAR = OFF_MISC + (EPI_GROUPS - 1) * 128 * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32 * 8 + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

enum { B_WFULL = 0, B_WEMPTY = 4, B_INFULL = 8, B_INEMPTY = 9, B_ACCA = 10, B_ACCB = 11, B_EPI = 12, B_B0DONE = 13, B_PGDONE = 14, B_COUNT = 15 };
constexpr uint8_t NOB = 0xff;
constexpr uint32_t TM_ACC_A = 0, TM_ACC_B = 256;

struct Seg {
  uint8_t wait_epi, wait_bar, commit0, commit1;
  uint8_t acc0, nfull, pad0, pad1;
  uint16_t N, kps, ktail, d_col;
  uint32_t a_off;
  uint32_t src_off;
};
constexpr int NUM_SEGS = 10;

struct Schedule {
  Seg s[NUM_SEGS];
  constexpr Schedule() : s{} {
    //        epi  bar        c0         c1          acc nf      N    kps ktail d_col     a_off    src
    s[0] = Seg{1, B_INFULL,  B_ACCA,    NOB,        0, 2, 0, 0, 256, 32, 16, TM_ACC_A, OFF_IN, (uint32_t)O_G0};   // G0
    s[1] = Seg{1, NOB,       B_ACCB,    NOB,        0, 8, 0, 0, 256, 32, 16, TM_ACC_B, OFF_X,  (uint32_t)O_G1};   // G1
    s[2] = Seg{1, NOB,       B_ACCA,    NOB,        0, 8, 0, 0, 256, 32, 16, TM_ACC_A, OFF_X,  (uint32_t)O_G2};   // G2
    s[3] = Seg{0, NOB,       B_ACCB,    NOB,        0, 8, 0, 0, 256, 32, 0,  TM_ACC_B, OFF_Y,  (uint32_t)O_B1};   // B1
    s[4] = Seg{1, NOB,       B_ACCA,    NOB,        0, 2, 0, 0, 256, 32, 16, TM_ACC_A, OFF_IN, (uint32_t)O_G0};   // G0'
    s[5] = Seg{1, NOB,       B_B0DONE,  NOB,        0, 4, 0, 0, 80,  64, 0,  TM_ACC_B, OFF_Y,  (uint32_t)O_B0};   // B0
    s[6] = Seg{0, NOB,       NOB,       NOB,        0, 1, 0, 0, 256, 32, 16, TM_ACC_A, OFF_IN, (uint32_t)O_C0};   // C0 (x, PE part)
    s[7] = Seg{0, NOB,       B_ACCA,    B_INEMPTY,  1, 8, 0, 0, 256, 32, 16, TM_ACC_A, OFF_X,  (uint32_t)(O_C0 + 256 * 48 * 2)};   // C0 (geo part)
    s[8] = Seg{1, B_PGDONE,  B_ACCB,    NOB,        0, 8, 0, 0, 256, 32, 16, TM_ACC_B, OFF_Y,  (uint32_t)O_C1};   // C1
    s[9] = Seg{1, NOB,       B_ACCA,    NOB,        0, 1, 0, 0, 16, 272, 0,  TM_ACC_A, OFF_X,  (uint32_t)O_C2};   // C2
  }
};
__constant__ Schedule c_sched = Schedule();

struct Params {
  const float* x; int64_t n;
  const uint8_t* blob; const float2* table; const float* scalings; int log2_T;
  float* sdf; float* grad; float* albedo;
  int64_t n_tiles;
  GridMode gm;      // nerfstudio torch grid (meta == nullptr) or an imported tiny-cuda-nn grid (nsk_common.cuh)
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

// softplus(beta=100) and its derivative sigmoid(100 z), from one exp
__device__ __forceinline__ float softplus_sig(float z, float& sig) {
  const float t = 100.0f * z;
  const float e = __expf(-fabsf(t));
  const float r = __fdividef(1.0f, 1.0f + e);
  sig = t >= 0.f ? r : e * r;
  return fmaxf(z, 0.f) + 0.01f * __logf(1.0f + e);
}
__device__ __forceinline__ float sig100(float z) {
  const float t = 100.0f * z;
  const float e = __expf(-fabsf(t));
  const float r = __fdividef(1.0f, 1.0f + e);
  return t >= 0.f ? r : e * r;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) sdf_field_tc_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32 * 8);
  float* sdf_part = reinterpret_cast<float*>(smem + OFF_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tailw = reinterpret_cast<const float*>(P.blob + STREAM_BYTES);   // W2[0,:] (256), b2[0]

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bars + 8 * (B_WFULL + i), 1); mbar_init(bars + 8 * (B_WEMPTY + i), 1); }
    mbar_init(bars + 8 * B_INFULL, PRO_THREADS / 32);      // one arrival per warp (lane 0 after __syncwarp): 512 per-thread arrivals on one
                                                           // mbarrier serialise in shared memory, seven times per tile
    mbar_init(bars + 8 * B_INEMPTY, 1);
    mbar_init(bars + 8 * B_ACCA, 1);
    mbar_init(bars + 8 * B_ACCB, 1);
    mbar_init(bars + 8 * B_EPI, EPI_THREADS / 32);
    mbar_init(bars + 8 * B_B0DONE, 1);
    mbar_init(bars + 8 * B_PGDONE, PRO_THREADS / 32);
    fence_barrier_init();
  }
  // constant-1 columns (k = 256, 257) of X and Y, written once
  if (threadIdx.x < 2 * TM) {
    const int r = threadIdx.x & (TM - 1);
    uint8_t* d = smem + (threadIdx.x < TM ? OFF_X : OFF_Y) + (uint32_t)(256 / 8) * (TM * 16) + r * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(d + TM * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
  reg_dec<48>();      // inside the role branch: ptxas compiles code after a join for the smallest budget reaching it
  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_last();
      uint32_t st = 0, ph = 0;
      for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int o = 0; o < NUM_SEGS; ++o) {
          const Seg sg = c_sched.s[o];
          const uint8_t* src = P.blob + sg.src_off;
          const int nst = sg.nfull + (sg.ktail ? 1 : 0);
#pragma unroll 1
          for (int i = 0; i < nst; ++i) {
            const uint32_t bytes = (uint32_t)sg.N * (uint32_t)(i < sg.nfull ? sg.kps : sg.ktail) * 2u;
            mbar_wait(bars + 8 * (B_WEMPTY + st), ph ^ 1);
            mbar_arrive_expect_tx(bars + 8 * (B_WFULL + st), bytes);
            bulk_g2s_hint(sbase + OFF_RING + st * STAGE_BYTES, src, bytes, bars + 8 * (B_WFULL + st), pol);
            src += bytes;
            if (++st == NSTAGE) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    uint32_t wst = 0, wph = 0, phases = 0;
    const uint64_t desc_hi = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32);   // version 1, SBO = 128 B
    bool first = true;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, first = false) {
#pragma unroll 1
      for (int o = 0; o < NUM_SEGS; ++o) {
        const Seg sg = c_sched.s[o];
        if (sg.wait_epi && !(first && o == 0)) {
          mbar_wait(bars + 8 * B_EPI, (phases >> B_EPI) & 1u);
          phases ^= 1u << B_EPI;
        }
        if (sg.wait_bar != NOB) {
          mbar_wait(bars + 8 * sg.wait_bar, (phases >> sg.wait_bar) & 1u);
          phases ^= 1u << sg.wait_bar;
        }
        tc_fence_after();
        const int N = sg.N;
        const int nst = sg.nfull + (sg.ktail ? 1 : 0);
        const uint32_t idesc = make_idesc_f16(TM, N);
        const uint32_t tmem_d = tmem + sg.d_col;
        const uint64_t ad0 = desc_hi | ((uint64_t)((TM * 16) >> 4) << 16) | (uint64_t)(((sbase + sg.a_off) >> 4) & 0x3FFF);
        const uint64_t bd0 = desc_hi | ((uint64_t)((N * 16) >> 4) << 16);
        const uint32_t b_step = (uint32_t)(2 * N * 16) >> 4;
        uint32_t kstep = 0, acc = sg.acc0;
#pragma unroll 1
        for (int i = 0; i < nst; ++i) {
          const int nmma = (i < sg.nfull ? sg.kps : sg.ktail) >> 4;
          mbar_wait(bars + 8 * (B_WFULL + wst), wph);
          tc_fence_after();
          if (elect_one()) {
            uint64_t ad = ad0 + (uint64_t)kstep * 256u;
            uint64_t bd = bd0 | (uint64_t)(((sbase + OFF_RING + wst * STAGE_BYTES) >> 4) & 0x3FFF);
            uint32_t a = acc;
#pragma unroll 1
            for (int j = 0; j < nmma; ++j) {
              umma_ss(tmem_d, ad, bd, idesc, a);
              a = 1;
              ad += 256;
              bd += b_step;
            }
            umma_commit(bars + 8 * (B_WEMPTY + wst));
          }
          __syncwarp();
          kstep += nmma;
          acc = 1;
          if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
        }
        if (elect_one()) {
          if (sg.commit0 != NOB) umma_commit(bars + 8 * sg.commit0);
          if (sg.commit1 != NOB) umma_commit(bars + 8 * sg.commit1);
        }
        __syncwarp();
      }
    }
  }
  } else if (warp >= EPI_WARP0 && warp < PRO_WARP0) {
    // ================================ epilogue ================================
    const int e = warp - EPI_WARP0;
    const int q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph_a = 0, ph_b = 0;
    uint8_t* const xrow = smem + OFF_X + (uint32_t)(hsel * (EPI_COLS / 8)) * (TM * 16) + row * 16;
    uint8_t* const yrow = smem + OFF_Y + (uint32_t)(hsel * (EPI_COLS / 8)) * (TM * 16) + row * 16;
    const uint32_t accA = tmem + TM_ACC_A + lane_off + hsel * EPI_COLS, accB = tmem + TM_ACC_B + lane_off + hsel * EPI_COLS;
    constexpr int NCB = EPI_COLS / 16;                 // 16-column blocks per thread

    auto wait_a = [&]() { mbar_wait(bars + 8 * B_ACCA, ph_a); ph_a ^= 1; tc_fence_after(); };
    auto wait_b = [&]() { mbar_wait(bars + 8 * B_ACCB, ph_b); ph_b ^= 1; tc_fence_after(); };
    auto done = [&]() { fence_proxy_async_smem(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bars + 8 * B_EPI); };
    // generic 128-column pass: f(values of 16 columns, column base) -> 8 packed words -> two 16-byte chunks of `dst`
    auto store16 = [&](uint8_t* dst, int cb, const uint32_t (&pk)[8]) {
      *reinterpret_cast<uint4*>(dst + (cb * 2) * (TM * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + (cb * 2 + 1) * (TM * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };

    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const int64_t sample = tile * TM + row;
      // ---- E1: a0 = softplus(z0) -> X --------------------------------------------------------------
      wait_a();
      {
        uint32_t v[2][16];
        tmem_ld16(accA, v[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) tmem_ld16(accA + (cb + 1) * 16, v[(cb + 1) & 1]);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float s0, s1;
            pk[j >> 1] = pack_h2(softplus_sig(__uint_as_float(v[cb & 1][j]), s0), softplus_sig(__uint_as_float(v[cb & 1][j + 1]), s1));
          }
          store16(xrow, cb, pk);
        }
      }
      done();
      // ---- E2: a1 -> X, g2 -> Y, sdf --------------------------------------------------------------------
      wait_b();
      {
        float sdf_acc = 0.f;
        uint32_t v[2][16];
        tmem_ld16(accB, v[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) tmem_ld16(accB + (cb + 1) * 16, v[(cb + 1) & 1]);
          uint32_t pa[8], pg[8];
          const float4* w4 = reinterpret_cast<const float4*>(tailw + hsel * EPI_COLS + cb * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w = __ldg(w4 + j4);
            const float ww[4] = {w.x, w.y, w.z, w.w};
            float a[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float sg;
              a[u] = softplus_sig(__uint_as_float(v[cb & 1][j4 * 4 + u]), sg);
              g[u] = sg * ww[u];
              sdf_acc = fmaf(a[u], ww[u], sdf_acc);
            }
            pa[j4 * 2] = pack_h2(a[0], a[1]); pa[j4 * 2 + 1] = pack_h2(a[2], a[3]);
            pg[j4 * 2] = pack_h2(g[0], g[1]); pg[j4 * 2 + 1] = pack_h2(g[2], g[3]);
          }
          store16(xrow, cb, pa);
          store16(yrow, cb, pg);
        }
        if (hsel > 0) sdf_part[(hsel - 1) * 128 + row] = sdf_acc;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (hsel == 0 && sample < P.n) {
          float tot = sdf_acc + __ldg(tailw + 256);
#pragma unroll
          for (int gsel = 0; gsel < EPI_GROUPS - 1; ++gsel) tot += sdf_part[gsel * 128 + row];
          P.sdf[sample] = tot;
        }
      }
      done();
      // ---- E3: geo feature -> X ---------------------------------------------------------------------------
      wait_a();
      {
        uint32_t v[2][16];
        tmem_ld16(accA, v[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) tmem_ld16(accA + (cb + 1) * 16, v[(cb + 1) & 1]);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_h2(__uint_as_float(v[cb & 1][j]), __uint_as_float(v[cb & 1][j + 1]));
          store16(xrow, cb, pk);
        }
      }
      done();
      // ---- E4: g1 = (W1^T g2) * sigmoid(100 z0) -> Y ------------------------------------------------------
      wait_b();
      wait_a();
      {
        uint32_t vb[2][16], va[2][16];
        tmem_ld16(accB, vb[0]); tmem_ld16(accA, va[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) { tmem_ld16(accB + (cb + 1) * 16, vb[(cb + 1) & 1]); tmem_ld16(accA + (cb + 1) * 16, va[(cb + 1) & 1]); }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            pk[j >> 1] = pack_h2(__uint_as_float(vb[cb & 1][j]) * sig100(__uint_as_float(va[cb & 1][j])),
                                 __uint_as_float(vb[cb & 1][j + 1]) * sig100(__uint_as_float(va[cb & 1][j + 1])));
          store16(yrow, cb, pk);
        }
      }
      done();
      // ---- E5: c0 = relu(.) -> Y  (C0's commit also covers B0, which read g1 from Y) ----------------------------
      wait_a();
      {
        uint32_t v[2][16];
        tmem_ld16(accA, v[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) tmem_ld16(accA + (cb + 1) * 16, v[(cb + 1) & 1]);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_h2(fmaxf(__uint_as_float(v[cb & 1][j]), 0.f), fmaxf(__uint_as_float(v[cb & 1][j + 1]), 0.f));
          store16(yrow, cb, pk);
        }
      }
      done();
      // ---- E6: c1 = relu(.) -> X ----------------------------------------------------------------------------
      wait_b();
      {
        uint32_t v[2][16];
        tmem_ld16(accB, v[0]);
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < NCB) tmem_ld16(accB + (cb + 1) * 16, v[(cb + 1) & 1]);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_h2(fmaxf(__uint_as_float(v[cb & 1][j]), 0.f), fmaxf(__uint_as_float(v[cb & 1][j + 1]), 0.f));
          store16(xrow, cb, pk);
        }
      }
      done();
      // ---- E7: albedo = sigmoid(.) ----------------------------------------------------------------------------
      wait_a();
      if (hsel == 0) {
        uint32_t v[8];
        tmem_ld8(tmem + TM_ACC_A + lane_off, v);
        tmem_ld_wait();
        if (sample < P.n) {
          P.albedo[sample * 3 + 0] = sigmoidf_(__uint_as_float(v[0]));
          P.albedo[sample * 3 + 1] = sigmoidf_(__uint_as_float(v[1]));
          P.albedo[sample * 3 + 2] = sigmoidf_(__uint_as_float(v[2]));
        }
      }
      done();
    }
  } else if (warp >= PRO_WARP0) {
    // ================================ prologue / gradient ================================
    reg_inc<112>();
    const int row = (warp - PRO_WARP0) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp - PRO_WARP0) * 32) << 16;
    const uint32_t mask = (1u << P.log2_T) - 1u;
    const float TWO_PI = 6.283185307179586f;
    uint32_t ph_empty = 0, ph_b0 = 0;

    // sin / cos of 2 pi x 2^f with the period removed exactly first (x 2^f is exact in fp32), then the fast SFU path
    auto pe_sincos = [&](float xd, int f, float& sn, float& cs) {
      float t = xd * (float)(1 << f);
      t -= rintf(t);
      __sincosf(TWO_PI * t, &sn, &cs);
    };
    // 4 hash levels per batch: 32 independent 8-byte gathers in flight per thread
    auto gather4 = [&](const float (&pos)[3], int lev0, float2 (&f)[4][8], float (&o)[4][3], float (&dwv)[4][3], float (&scv)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t idx[8];
        grid_corners(P.gm, lev0 + u, pos[0], pos[1], pos[2], __ldg(P.scalings + lev0 + u), mask, idx, o[u][0], o[u][1], o[u][2], dwv[u], scv[u]);
        const float2* tl = P.table + ((size_t)(lev0 + u) << P.log2_T);
#pragma unroll
        for (int c = 0; c < 8; ++c) f[u][c] = __ldg(tl + idx[c]);
      }
    };

    // forward features of one sample -> 40 packed words (the IN row):
    // [x_hi 3 | PE sin 18 | PE cos 18 | feat 32 | 1, 1 | x_lo 3 | 0 x4]
    auto features = [&](int64_t sample, uint32_t (&w)[40]) {
      const int64_t s = min(sample, P.n - 1);
      const float xv[3] = {__ldg(P.x + s * 3), __ldg(P.x + s * 3 + 1), __ldg(P.x + s * 3 + 2)};
      float pos[3], J[9];
      sdf_contract(xv, pos, J);
      float h[40];   // columns 0..39 (x, PE, feat0) assembled as floats, the rest packed on the fly
      float xl[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float hi = __half2float(__float2half_rn(xv[d]));
        h[d] = hi;
        xl[d] = xv[d] - hi;
#pragma unroll
        for (int f = 0; f < 6; ++f) pe_sincos(xv[d], f, h[3 + d * 6 + f], h[3 + 18 + d * 6 + f]);
      }
      float carry = 0.f;   // feature waiting for its pair partner (features start at the odd column 39)
#pragma unroll
      for (int lev0 = 0; lev0 < SDF_LEVELS; lev0 += 4) {
        float2 f[4][8];
        float o[4][3], dwv[4][3], scv[4];
        gather4(pos, lev0, f, o, dwv, scv);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 r = hash_interp(f[u], o[u][0], o[u][1], o[u][2]);
          const int lev = lev0 + u;
          if (lev == 0) h[39] = r.x;
          else w[19 + lev] = pack_h2(carry, r.x);       // columns 38+2lev, 39+2lev
          carry = r.y;                                    // column 40+2lev
        }
      }
#pragma unroll
      for (int i = 0; i < 20; ++i) w[i] = pack_h2(h[2 * i], h[2 * i + 1]);
      w[35] = pack_h2(carry, 1.0f);          // columns 70 (feat 31), 71 (bias hi)
      w[36] = pack_h2(1.0f, xl[0]);          // columns 72 (bias lo), 73
      w[37] = pack_h2(xl[1], xl[2]);         // columns 74, 75
      w[38] = 0u; w[39] = 0u;
    };
    auto write_in = [&](const uint32_t (&w)[40]) {
      uint8_t* d = smem + OFF_IN + row * 16;
#pragma unroll
      for (int kc = 0; kc < 10; ++kc) *reinterpret_cast<uint4*>(d + kc * (TM * 16)) = make_uint4(w[kc * 4], w[kc * 4 + 1], w[kc * 4 + 2], w[kc * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_INFULL);
    };

    uint32_t w[40];
    if ((int64_t)blockIdx.x < P.n_tiles) {
      features((int64_t)blockIdx.x * TM + row, w);
      write_in(w);
    }
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const int64_t next = tile + gridDim.x;
      if (next < P.n_tiles) features(next * TM + row, w);       // overlaps the MMA chain of `tile`
      // ---- g0 = d sdf / d(layer-0 input) of this tile, straight from TMEM (B0) -----------------------------------
      mbar_wait(bars + 8 * B_B0DONE, ph_b0); ph_b0 ^= 1;
      tc_fence_after();
      const int64_t sample = tile * TM + row;
      const int64_t s = min(sample, P.n - 1);
      const float xv[3] = {__ldg(P.x + s * 3), __ldg(P.x + s * 3 + 1), __ldg(P.x + s * 3 + 2)};
      float gx[3];
      {
        // columns 0..39: x (3), PE sin (18), PE cos (18), pad ; hash features start at column 40
        uint32_t g[5][8];
#pragma unroll
        for (int i = 0; i < 5; ++i) tmem_ld8(tmem + TM_ACC_B + lane_off + i * 8, g[i]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float acc = __uint_as_float(g[0][d]);
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            float sn, cs;
            pe_sincos(xv[d], f, sn, cs);
            const int cs_i = 3 + d * 6 + f, cc_i = 3 + 18 + d * 6 + f;
            acc += TWO_PI * (float)(1 << f) * (cs * __uint_as_float(g[cs_i >> 3][cs_i & 7]) - sn * __uint_as_float(g[cc_i >> 3][cc_i & 7]));
          }
          gx[d] = acc;
        }
      }
      uint32_t gfe[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) tmem_ld8(tmem + TM_ACC_B + lane_off + 40 + i * 8, gfe[i]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_PGDONE);     // the TMEM columns of B0 may be overwritten (C1)
      // ---- hand the next tile's inputs over as soon as C0 of this tile has consumed IN ------------------------------
      if (next < P.n_tiles) {
        mbar_wait(bars + 8 * B_INEMPTY, ph_empty); ph_empty ^= 1;
        write_in(w);
      }
      // ---- chain rule through the trilinear hash interpolation and the contraction (overlaps the next tile's MMAs) ----
      {
        float pos[3], J[9];
        sdf_contract(xv, pos, J);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int lev0 = 0; lev0 < SDF_LEVELS; lev0 += 4) {
          float2 f[4][8];
          float o[4][3], dwv[4][3], scv[4];
          gather4(pos, lev0, f, o, dwv, scv);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int lev = lev0 + u;
            const float sc = scv[u];
            float fa[8], fb[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { fa[c] = f[u][c].x; fb[c] = f[u][c].y; }
            float da[3], db[3];
            hash_interp_grad(fa, o[u][0], o[u][1], o[u][2], da);
            hash_interp_grad(fb, o[u][0], o[u][1], o[u][2], db);
            const float ga = __uint_as_float(gfe[(2 * lev) >> 3][(2 * lev) & 7]), gb = __uint_as_float(gfe[(2 * lev + 1) >> 3][(2 * lev + 1) & 7]);
            a0 += sc * dwv[u][0] * (da[0] * ga + db[0] * gb);
            a1 += sc * dwv[u][1] * (da[1] * ga + db[1] * gb);
            a2 += sc * dwv[u][2] * (da[2] * ga + db[2] * gb);
          }
        }
        gx[0] += J[0] * a0 + J[3] * a1 + J[6] * a2;
        gx[1] += J[1] * a0 + J[4] * a1 + J[7] * a2;
        gx[2] += J[2] * a0 + J[5] * a1 + J[8] * a2;
      }
      if (sample < P.n) { P.grad[sample * 3] = gx[0]; P.grad[sample * 3 + 1] = gx[1]; P.grad[sample * 3 + 2] = gx[2]; }
    }
  }


  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace sdftc
}  // namespace nsk

extern "C" int64_t nsk_sdf_tc_weights_bytes(void) { return nsk::sdftc::BLOB_BYTES; }

extern "C" int nsk_sdf_field_tc_fwd(const float* x, int64_t n, const void* sdf_weights, const float* hash_table,
                                    const float* scalings, int num_levels, int log2_T, float* sdf, float* grad,
                                    float* albedo, void* stream) {
  return nsk_sdf_field_tc_fwd_ex(x, n, sdf_weights, hash_table, scalings, num_levels, log2_T, nullptr, 0, sdf, grad, albedo, stream);
}

extern "C" int nsk_sdf_field_tc_fwd_ex(const float* x, int64_t n, const void* sdf_weights, const float* hash_table,
                                       const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                                       float* sdf, float* grad, float* albedo, void* stream) {
  using namespace nsk::sdftc;
  NSK_REQUIRE(grid_meta == nullptr || (reinterpret_cast<uintptr_t>(grid_meta) & 15) == 0, "nsk_sdf_field_tc_fwd_ex: grid_meta must be 16-byte aligned");
  NSK_REQUIRE(num_levels == nsk::SDF_LEVELS, "nsk_sdf_field_tc_fwd: the SDF position encoding has 16 levels");
  if (n == 0) return 0;
  NSK_REQUIRE(x && sdf_weights && hash_table && scalings && sdf && grad && albedo, "nsk_sdf_field_tc_fwd: null pointer");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(sdf_weights) & 15) == 0, "nsk_sdf_field_tc_fwd: weight blob must be 16-byte aligned");
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "nsk_sdf_field_tc_fwd: device setup", &num_sms, [] {
        return cudaFuncSetAttribute(sdf_field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      }))
    return err;
  Params P;
  P.x = x; P.n = n; P.blob = reinterpret_cast<const uint8_t*>(sdf_weights);
  P.table = reinterpret_cast<const float2*>(hash_table); P.scalings = scalings; P.log2_T = log2_T;
  P.sdf = sdf; P.grad = grad; P.albedo = albedo;
  P.gm = nsk::GridMode{reinterpret_cast<const int4*>(grid_meta), smoothstep};
  P.n_tiles = (n + TM - 1) / TM;
  const int64_t grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  sdf_field_tc_kernel<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  return nsk::check_launch("sdf_field_tc_kernel");
}

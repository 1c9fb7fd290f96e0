// Training-path dense contractions on the 5th-gen tensor cores: tcgen05.mma.kind::tf32 with fp32 operands read
// straight from HBM (no fp16 copies of activations or gradients), fp32 accumulators in TMEM.
//
//   nsk_gemm_tf32_nt : C[M,N]  = act(A[M,K] . B[N,K]^T + bias[N]) (+ C)      forward layers and dX = dY . W
//   nsk_gemm_tf32_tn : C[P,Q] += sum_m A[m,P]^T . B[m,Q]                      dW = dY^T . X, split over m across CTAs
//
// These are the layer contractions of the reference's torch autograd graph for the SDF/colour MLP
// (neusky/fields/sdf_albedo_field.py:185-269) and the FiLM-SIREN DDF (ns_reni/reni/field_components/film_siren.py:45-156)
// in the training step; the fused forward-only kernels (sdf_field_tc.cu, sky_shade_tc2.cu) stay the eval path.
//
// One persistent CTA per SM, 544 threads, warp-specialised:
//   warps 0-7   operand staging.  3xTF32 NT (the training step's forward layers and dX): one thread issues tensor-map TMA boxes
//               (raw fp32 = the hi plane, the MMA truncates to 19 bits), warps 4-7 derive the lo plane (v - trunc(v)) for the
//               3xTF32 scheme (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo), whose result is fp32-accurate and is what the parity tests
//               pin.  3xTF32 TN (dW): the same roles with MN-major tensor-map boxes (both operands are read as they lie in HBM,
//               rows = reduction index).  Other variants / fallbacks: LDGSTS / ld.global -> registers -> st.shared in the
//               no-swizzle K-major core-matrix layout [K/4][rows][4 x tf32] (TN transposes in registers).
//   warp 16     one lane issues tcgen05.mma (M=128, N<=256, K=8 per instruction) into a double-buffered TMEM accumulator
//   warps 8-15  epilogue (two per TMEM lane quadrant, alternating 32-column chunks): tcgen05.ld, bias + activation,
//               st.global (NT) / red.global.add (TN).  The epilogue, not the tensor pipe or HBM, was what bounded the first
//               version (4 warps, ~3300 cycles per 32x32 chunk: profiles/r01_gemm_tf32_ncu_before.txt).
// Stage ring: NT 3 x 48 KB, TN 4 x 48 KB (2 x 96 KB with split=3), mbarrier full/empty; accumulator ring: 2 x 256 TMEM columns.
#include <algorithm>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#include "nsk_common.cuh"
#include "tc_util.cuh"

namespace nsk {
namespace gemm {
using namespace nsk::tc;

constexpr int TM = 128;
constexpr int BN = 256;
constexpr int PROD_WARPS = 8;
constexpr int EPI_WARP0 = 8;
constexpr int EPI_WARPS = 8;                                // two per TMEM lane quadrant, interleaved 32-column chunks
constexpr int MMA_WARP = 16;
constexpr int THREADS = 17 * 32;
constexpr uint32_t A_LBO = TM * 16, B_LBO = BN * 16;
// Pipeline shape per variant.  The 3xTF32 NT variant (every forward layer and every dX of the training step) is the "pair"
// shape: a work item is TWO 128-row tiles of A against one <=256-row tile of B, so every B chunk fetched from L2 feeds twice
// the MMA work (B is 2/3 of the load traffic with single tiles); chunks are 16 wide in a 3-deep LDGSTS ring (with hi + lo
// planes a stage costs twice its global bytes in shared memory).  The two tiles use both TMEM accumulators, so the epilogue
// of an item does not overlap the next item's MMAs; it is short since it runs on 8 warps.  With LDGSTS the variant was bound
// by the issue rate of the 4 loader warps (ncu: loaders stalled issuing, epilogue warps 2/3 idle); the operands therefore
// arrive as tensor-map TMA boxes (TMA = true: SWIZZLE_64B, one issuing thread), the LDGSTS loader stays as the fallback.
template <int SPLIT, bool TN, bool TMA = false> struct Cfg {
  static constexpr bool PAIR = !TN && (SPLIT == 3 || TMA);                     // plain tf32 takes the pair shape only with TMA operands
  static constexpr bool TNT = TN && TMA;                                       // TN fed by MN-major tensor-map boxes
  static constexpr bool ROLES = PAIR || TNT;                                   // loader (+ hi/lo splitter warps when SPLIT == 3)
  static constexpr int KC = (PAIR || TNT) ? 16 : 32;
  // NT keeps 18 KB for the epilogue staging.  Plain tf32 (SPLIT == 1) has no lo planes: the same 192 KB hold twice the stages,
  // and the MMA issuer waits on the TMA (RAW) barrier directly -- the variant is HBM-bound, ring depth is what it needs.
  static constexpr int ST = PAIR ? (SPLIT == 3 ? 3 : 6) : (TNT ? (SPLIT == 3 ? 4 : 8) : (TN ? (SPLIT == 3 ? 2 : 4) : 3));
  static constexpr int TMI = PAIR ? 2 * TM : TM;                              // rows of the A operand per work item
  static constexpr uint32_t A_HALF = TM * KC * 4;
  static constexpr uint32_t A_BYTES = TMI * KC * 4;
  static constexpr uint32_t B_BYTES = BN * KC * 4;
  static constexpr uint32_t LO_OFF = A_BYTES + B_BYTES;                       // split 3: lo planes follow the hi planes
  static constexpr uint32_t STAGE = (A_BYTES + B_BYTES) * (SPLIT == 3 ? 2 : 1);
};
constexpr int KC_TN = 32;   // host-side rounding of the TN row split
constexpr int EPI_LD = 36;                                  // floats per staged row (32 + 4: conflict-free 16-byte accesses)
constexpr int EPI_ROWS = 16;                                // rows transposed per pass (two passes per 32 x 32 chunk)
constexpr uint32_t EPI_BYTES = EPI_WARPS * EPI_ROWS * EPI_LD * 4;   // one 16 x 32 transposition tile per epilogue warp (NT only)

struct Params {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int64_t M;        // NT: rows of A and C.  TN: reduction length (rows of A and B)
  int N;            // NT: columns of C (rows of B).  TN: P = columns of A = rows of C
  int K;            // NT: reduction length.  TN: Q = columns of B = columns of C
  int lda, ldb, ldc;
  const float* aux; // NT only: forward OUTPUT of the activation whose derivative multiplies the result (dact != 0)
  int ldaux;
  int act;          // NT only: 0 none, 1 relu, 2 leaky relu 0.2, 3 softplus beta=100 (torch threshold 20), 4 sigmoid
  int dact;         // NT only: result *= act'(.) expressed through the activation's output aux[row, col] (same codes)
  int accumulate;   // NT only: C += result
  int n_btiles;     // tiles along the B-operand rows
  int n_atiles;     // TN only: tiles along P
  int64_t rows_per_split;  // TN only
  int64_t n_items;
};

struct Item {
  int64_t a0;   // first A-operand row (NT: row of A; TN: column of A)
  int b0;       // first B-operand row (NT: row of B; TN: column of B)
  int bn;       // MMA N: valid B-operand rows rounded up to 16
  int64_t r0, r1;  // reduction range
};

template <bool TN, int TMI = TM>
__device__ __forceinline__ Item get_item(const Params& p, int64_t w) {
  Item it;
  if (!TN) {
    const int64_t mt = w / p.n_btiles;
    const int nt = (int)(w % p.n_btiles);
    it.a0 = mt * TMI;
    it.b0 = nt * BN;
    it.bn = (min(BN, p.N - it.b0) + 15) & ~15;
    it.r0 = 0;
    it.r1 = p.K;
  } else {
    const int per = p.n_atiles * p.n_btiles;
    const int64_t split = w / per;
    const int rem = (int)(w % per);
    it.a0 = (int64_t)(rem / p.n_btiles) * TM;
    it.b0 = (rem % p.n_btiles) * BN;
    it.bn = (min(BN, p.K - it.b0) + 15) & ~15;
    it.r0 = split * p.rows_per_split;
    it.r1 = min(p.M, it.r0 + p.rows_per_split);
  }
  return it;
}

// Shared-memory descriptor for a K-major operand written by TMA with CU_TENSOR_MAP_SWIZZLE_64B: rows of 64 B (16 tf32), 8-row
// atoms of 512 B (SBO), 16-byte chunks XOR-swizzled by address bits [7,9); layout type 4 = SWIZZLE_64B; LBO unused (one k-step
// of 8 tf32 = 32 B stays inside the swizzle span).  k-steps advance the start address by 32 B.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// MN-major fp32/tf32 operand written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B as [MN block of 32][k rows][128 B]: 32 tf32
// of the MN dimension are contiguous (128 B), consecutive k rows are 128 B apart, 32-byte chunks are XOR-swizzled over 4 rows
// (Swizzle<2,5,2>: the atom is 32 MN x 4 K).  UMMA layout type 1 = SWIZZLE_128B_BASE32B, which is the mode 32-bit MN-major
// operands need (with the plain 128B/16B-atom mode the MMA returns zeros; bring-up notes: scripts/tn_debug.py).
// LBO = byte stride between 32-wide MN blocks, SBO = stride between the two 4-row k atoms of one K=8 instruction (512 B);
// k-steps advance the start address by 8 rows = 1024 B.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}

// 2-D tiled TMA load global -> shared, completion (bytes) on an mbarrier.  c0 = innermost coordinate (k), c1 = row.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// Start of the c-th reduction chunk of a work item.  NT: every CTA walks the same [N,K] weight matrix, and CTAs launched
// together stay in lockstep, so with a common order all 148 SMs ask the same few L2 slices for the same 16-32 KB chunk at
// the same moment (~1 us per chunk whatever N or the split mode, measured).  Rotating the chunk order by the CTA index
// spreads the concurrent requests over the whole matrix.  Producer and MMA issuer must use the same order (the k tail).
template <bool TN>
__device__ __forceinline__ int64_t chunk_start(const Item& it, int c, int nch, int kc) {
  if (TN) return it.r0 + (int64_t)c * kc;
  return it.r0 + (int64_t)((c + (int)(blockIdx.x % (unsigned)nch)) % nch) * kc;
}

// fp32 accumulate, tf32 A and B, both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool mn_major = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// LDGSTS with zero fill (src-size 0) for out-of-range elements; completion is observed through cp_async_arrive
__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4u : 0u) : "memory");
}
// the mbarrier receives one arrival from this thread once all its prior cp.async have landed (.noinc: the arrival is part
// of the barrier's expected count)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int SPLIT, uint32_t LO_OFF>
__device__ __forceinline__ void put(uint8_t* hi_plane, uint32_t off, float4 v) {
  if (SPLIT == 1) {
    *reinterpret_cast<float4*>(hi_plane + off) = v;
  } else {
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4*>(hi_plane + off) = h;
    *reinterpret_cast<float4*>(hi_plane + LO_OFF + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}
// lo plane from a raw fp32 piece already sitting in the hi plane: tcgen05.mma.kind::tf32 truncates its operands to the top 19
// bits (profiles/r01_tf32_truncation_probe.log), so the raw tile IS the hi operand and only v - trunc(v) has to be written
template <uint32_t LO_OFF>
__device__ __forceinline__ void fix_lo(uint8_t* hi_plane, uint32_t off) {
  const float4 v = *reinterpret_cast<const float4*>(hi_plane + off);
  *reinterpret_cast<float4*>(hi_plane + LO_OFF + off) =
      make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return v > 0.0f ? v : 0.2f * v;
  if (act == 3) return (v * 100.0f > 20.0f) ? v : log1pf(expf(v * 100.0f)) * 0.01f;
  if (act == 4) return 1.0f / (1.0f + expf(-v));
  return v;
}
template <int ACT>
__device__ __forceinline__ float4 bias_act4(float4 t, float4 b) {
  return make_float4(act_apply(t.x + b.x, ACT), act_apply(t.y + b.y, ACT), act_apply(t.z + b.z, ACT), act_apply(t.w + b.w, ACT));
}
// derivative of activation `dact` written in terms of its output a
__device__ __forceinline__ float dact_from_output(float a, int dact) {
  if (dact == 1) return a > 0.0f ? 1.0f : 0.0f;
  if (dact == 2) return a > 0.0f ? 1.0f : 0.2f;
  if (dact == 3) return -expm1f(-100.0f * a);        // softplus_100: sigmoid(100 z) = 1 - exp(-100 a)
  if (dact == 4) return a * (1.0f - a);
  return 1.0f;
}

template <int SPLIT, bool TN, bool TMA>
__global__ void __launch_bounds__(THREADS, 1) gemm_tf32_kernel(const Params p, const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB) {
  using C = Cfg<SPLIT, TN, TMA>;
  constexpr int ST = C::ST, KC = C::KC, TMI = C::TMI;
  constexpr bool PAIR = C::PAIR, TNT = C::TNT, ROLES = C::ROLES;
  constexpr uint32_t STAGE = C::STAGE, A_BYTES = C::A_BYTES, B_BYTES = C::B_BYTES, LO_OFF = C::LO_OFF, A_HALF = C::A_HALF;
  (void)A_HALF;
  (void)B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* const bars_p = reinterpret_cast<uint64_t*>(smem + ST * STAGE);
  const uint32_t bars = smem_u32(bars_p);
  const uint32_t FULL = bars, EMPTY = bars + 8 * ST, ACCF = bars + 16 * ST, ACCE = bars + 16 * ST + 16, RAW = bars + 16 * ST + 32;
  (void)RAW;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + ST * STAGE + 248);   // barriers: (3 ST + 4) x 8 B <= 152 B

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(FULL + 8 * s, ROLES ? 4 : PROD_WARPS * 32);      // ROLES: one arrival per splitter warp (lane 0 after __syncwarp)
      mbar_init(EMPTY + 8 * s, 1);
      if (ROLES) mbar_init(RAW + 8 * s, TMA ? 1 : 128);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(ACCF + 8 * b, 1);
      mbar_init(ACCE + 8 * b, EPI_WARPS);                        // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < PROD_WARPS) {
    // ------------------------------------------------------------------ operand staging
    uint32_t stage = 0, phase = 0;
    const int r8 = lane & 7, kq_lo = lane >> 3;
    if constexpr (TNT) {
      // dW = dY^T X with both operands read as they lie in HBM (rows = reduction index m): MN-major tensor-map boxes
      // [32 columns x KC rows x 4 | 8 column blocks], SWIZZLE_128B_ATOM_32B; no register transposition, no L1tex traffic.
      if (warp == 0 && lane == 0) {
        for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
          const Item it = get_item<TN, TMI>(p, w);
          for (int64_t r = it.r0; r < it.r1; r += KC) {
            mbar_wait(EMPTY + 8 * stage, phase ^ 1);
            const uint32_t sa32 = smem_u32(smem + stage * STAGE);
            mbar_arrive_expect_tx(RAW + 8 * stage, A_BYTES + B_BYTES);
            tma_load_3d(sa32, &tmA, 0, (int)r, (int)(it.a0 >> 5), RAW + 8 * stage);
            tma_load_3d(sa32 + A_BYTES, &tmB, 0, (int)r, it.b0 >> 5, RAW + 8 * stage);
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
        }
      } else if (SPLIT == 3 && warp >= 4) {
        const uint32_t t128 = (uint32_t)(threadIdx.x & 127) * 16;
        for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
          const Item it = get_item<TN, TMI>(p, w);
          for (int64_t r = it.r0; r < it.r1; r += KC) {
            mbar_wait(RAW + 8 * stage, phase);
            uint8_t* const sa = smem + stage * STAGE;
#pragma unroll
            for (int j = 0; j < (int)((A_BYTES + B_BYTES) / 2048); ++j) fix_lo<LO_OFF>(sa, t128 + (uint32_t)j * 2048);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(FULL + 8 * stage);
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if constexpr (PAIR) {
      // Warps 0-3 LOAD: raw fp32 chunks go global -> shared with LDGSTS (they are the hi planes as they are: the MMA
      // truncates); completion is counted on the stage's RAW mbarrier, so the whole ring is in flight.  Warps 4-7 SPLIT: wait
      // RAW, derive the lo planes, publish FULL.  Two roles because fence.proxy.async -- needed before the MMA may read the lo
      // planes -- also waits for the issuing thread's outstanding LDGSTS: with one role doing both, every chunk paid a full
      // global-load latency (~1 us per chunk whatever the ring depth, measured).
      const int spi = (p.K + KC - 1) / KC;                                        // chunks per work item
      const int rot = (int)(blockIdx.x % spi);                                    // see chunk_start()
      // warp item = 8 rows x 4 k-quads (64 B per row): 8 distinct 128-byte lines per LDGSTS instruction (the L1tex pipe spends
      // ~2 cycles per line touched, which is what bounded the 8-wide chunk variant: 16 lines per instruction) and 8 distinct
      // 16-byte bank groups per quarter warp on the shared-memory side
      constexpr int NP = 8;                                                        // pieces of A (and of B) per thread per chunk
      const int kq = lane >> 3;
      const int w4 = warp & 3;
      int rowv[NP];
      uint32_t offA[NP], offB[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        rowv[j] = (w4 + 4 * j) * 8 + r8;                                          // 0..255
        offA[j] = (uint32_t)((rowv[j] >> 7) * A_HALF + kq * A_LBO + (rowv[j] & 127) * 16);
        offB[j] = (uint32_t)(kq * B_LBO + rowv[j] * 16);
      }
      if (TMA && warp < 4) {
        // one thread feeds the ring with two tensor-map boxes per chunk (A: 256 rows x 64 B, B: 256 rows x 64 B, SWIZZLE_64B,
        // out-of-range rows / columns zero-filled by the TMA unit): no L1tex wavefronts, no per-thread address math
        if (warp == 0 && lane == 0) {
          for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
            const Item it = get_item<TN, TMI>(p, w);
            for (int c = 0; c < spi; ++c) {
              int cc = c + rot;
              if (cc >= spi) cc -= spi;
              mbar_wait(EMPTY + 8 * stage, phase ^ 1);
              const uint32_t sa32 = smem_u32(smem + stage * STAGE);
              mbar_arrive_expect_tx(RAW + 8 * stage, A_BYTES + B_BYTES);
              tma_load_2d(sa32, &tmA, cc * KC, (int)it.a0, RAW + 8 * stage);
              tma_load_2d(sa32 + A_BYTES, &tmB, cc * KC, it.b0, RAW + 8 * stage);
              if (++stage == ST) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else if (warp < 4) {
        for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
          const Item it = get_item<TN, TMI>(p, w);
          const float* ap[NP];
          const float* bp[NP];
          bool aok[NP], bok[NP];
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const int64_t grow = it.a0 + rowv[j];
            const int n = it.b0 + rowv[j];
            aok[j] = grow < p.M;
            bok[j] = rowv[j] < it.bn && n < p.N;
            ap[j] = aok[j] ? p.A + grow * p.lda + kq * 4 : p.A;
            bp[j] = bok[j] ? p.B + (int64_t)n * p.ldb + kq * 4 : p.B;
          }
          for (int c = 0; c < spi; ++c) {
            int cc = c + rot;
            if (cc >= spi) cc -= spi;
            const int k0 = cc * KC;
            const bool kok = k0 + kq * 4 < it.r1;
            mbar_wait(EMPTY + 8 * stage, phase ^ 1);
            const uint32_t sa32 = smem_u32(smem + stage * STAGE), sb32 = sa32 + A_BYTES;
#pragma unroll
            for (int j = 0; j < NP; ++j) cp_async16(sa32 + offA[j], (aok[j] && kok) ? ap[j] + k0 : p.A, aok[j] && kok);
#pragma unroll
            for (int j = 0; j < NP; ++j) cp_async16(sb32 + offB[j], (bok[j] && kok) ? bp[j] + k0 : p.B, bok[j] && kok);
            cp_async_arrive(RAW + 8 * stage);
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
        }
      } else if (SPLIT == 3) {
        for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
          for (int c = 0; c < spi; ++c) {
            mbar_wait(RAW + 8 * stage, phase);
            uint8_t* const sa = smem + stage * STAGE;
            uint8_t* const sb = sa + A_BYTES;
            if (TMA) {                                       // layout-agnostic: hi and lo planes share the (swizzled) layout
              const uint32_t t128 = (uint32_t)(threadIdx.x & 127) * 16;
#pragma unroll
              for (int j = 0; j < 2 * NP; ++j) fix_lo<LO_OFF>(sa, t128 + (uint32_t)j * 2048);
            } else {
#pragma unroll
              for (int j = 0; j < NP; ++j) fix_lo<LO_OFF>(sa, offA[j]);
#pragma unroll
              for (int j = 0; j < NP; ++j) fix_lo<LO_OFF>(sb, offB[j]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(FULL + 8 * stage);
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else
    for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x) {
      const Item it = get_item<TN, TMI>(p, w);
      const int nch = (int)((it.r1 - it.r0 + KC - 1) / KC);
      for (int c = 0; c < nch; ++c) {
        const int64_t r = chunk_start<TN>(it, c, nch, KC);
        mbar_wait(EMPTY + 8 * stage, phase ^ 1);
        uint8_t* const sa = smem + stage * STAGE;
        uint8_t* const sb = sa + A_BYTES;
        if (SPLIT == 1) {
          // asynchronous path: LDGSTS straight into the operand layout, completion counted on the stage's mbarrier, so
          // every stage of the ring is in flight at once (the register path below has one chunk in flight per thread)
          const uint32_t sa32 = smem_u32(sa), sb32 = sa32 + A_BYTES;
          if (!TN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
              const int64_t grow = it.a0 + row, k = r + kq * 4;
              const bool ok = grow < p.M && k < it.r1;
              cp_async16(sa32 + kq * A_LBO + row * 16, ok ? p.A + grow * p.lda + k : p.A, ok);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
              const int64_t k = r + kq * 4;
              const int n = it.b0 + row;
              const bool ok = row < it.bn && n < p.N && k < it.r1;
              cp_async16(sb32 + kq * B_LBO + row * 16, ok ? p.B + (int64_t)n * p.ldb + k : p.B, ok);
            }
          } else {
            // lane = (m & 3) + 4 * (n & 7): a warp instruction reads 4 rows x one 32-byte sector and writes 128 contiguous bytes
            const int ml = lane & 3, nl = lane >> 2;
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {           // A operand: 128 n x 32 m = 16 n-groups x 8 m-quads = 128 warp items
              const int wi = warp + 8 * j, n = (wi & 15) * 8 + nl, mq = wi >> 4;
              const int64_t col = it.a0 + n, m = r + mq * 4 + ml;
              const bool ok = col < p.N && m < it.r1;
              cp_async4(sa32 + mq * A_LBO + n * 16 + ml * 4, ok ? p.A + m * p.lda + col : p.A, ok);
            }
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {           // B operand: 256 n x 32 m = 32 n-groups x 8 m-quads = 256 warp items
              const int wi = warp + 8 * j, n = (wi & 31) * 8 + nl, mq = wi >> 5;
              const int64_t m = r + mq * 4 + ml;
              const int col = it.b0 + n;
              const bool ok = n < it.bn && col < p.K && m < it.r1;
              cp_async4(sb32 + mq * B_LBO + n * 16 + ml * 4, ok ? p.B + m * p.ldb + col : p.B, ok);
            }
          }
          cp_async_arrive(FULL + 8 * stage);
          if (++stage == ST) { stage = 0; phase ^= 1; }
          continue;
        }
        float4 va[4], vb[8];
        if (!TN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
            const int64_t grow = it.a0 + row, k = r + kq * 4;
            va[j] = (grow < p.M && k < it.r1) ? ldg4(p.A + grow * p.lda + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
            const int64_t k = r + kq * 4;
            const int n = it.b0 + row;
            vb[j] = (row < it.bn && n < p.N && k < it.r1) ? ldg4(p.B + (int64_t)n * p.ldb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
            put<SPLIT, LO_OFF>(sa, (uint32_t)(kq * A_LBO + row * 16), va[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int wi = warp + 8 * j, row = (wi >> 1) * 8 + r8, kq = (wi & 1) * 4 + kq_lo;
            put<SPLIT, LO_OFF>(sb, (uint32_t)(kq * B_LBO + row * 16), vb[j]);
          }
        } else {
          // element (operand row n, k = m): X[(r + k) * ld + col0 + n]; 4 consecutive m per thread -> one 16-byte store
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int wi = warp + 8 * j, n = (wi & 3) * 32 + lane, mq = wi >> 2;
            const int64_t col = it.a0 + n, m = r + mq * 4;
            const bool cv = col < p.N;
            const float* src = p.A + m * p.lda + col;
            va[j].x = (cv && m + 0 < it.r1) ? __ldg(src) : 0.f;
            va[j].y = (cv && m + 1 < it.r1) ? __ldg(src + p.lda) : 0.f;
            va[j].z = (cv && m + 2 < it.r1) ? __ldg(src + 2 * (int64_t)p.lda) : 0.f;
            va[j].w = (cv && m + 3 < it.r1) ? __ldg(src + 3 * (int64_t)p.lda) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int wi = warp + 8 * j, n = (wi & 7) * 32 + lane, mq = wi >> 3;
            const int64_t m = r + mq * 4;
            const int col = it.b0 + n;
            const bool cv = n < it.bn && col < p.K;
            const float* src = p.B + m * p.ldb + col;
            vb[j].x = (cv && m + 0 < it.r1) ? __ldg(src) : 0.f;
            vb[j].y = (cv && m + 1 < it.r1) ? __ldg(src + p.ldb) : 0.f;
            vb[j].z = (cv && m + 2 < it.r1) ? __ldg(src + 2 * (int64_t)p.ldb) : 0.f;
            vb[j].w = (cv && m + 3 < it.r1) ? __ldg(src + 3 * (int64_t)p.ldb) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int wi = warp + 8 * j, n = (wi & 3) * 32 + lane, mq = wi >> 2;
            put<SPLIT, LO_OFF>(sa, (uint32_t)(mq * A_LBO + n * 16), va[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int wi = warp + 8 * j, n = (wi & 7) * 32 + lane, mq = wi >> 3;
            put<SPLIT, LO_OFF>(sb, (uint32_t)(mq * B_LBO + n * 16), vb[j]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(FULL + 8 * stage);
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issue (one lane)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x, ++iter) {
        const Item it = get_item<TN, TMI>(p, w);
        if constexpr (PAIR) {
          // both accumulators belong to this item (tile halves 0 / 1); each barrier completes once per item
          mbar_wait(ACCE, (iter & 1) ^ 1);
          mbar_wait(ACCE + 8, (iter & 1) ^ 1);
          tc_fence_after();
          const uint32_t idesc = make_idesc_tf32(TM, it.bn);
          const int halves = (it.a0 + TM < p.M) ? 2 : 1;
          const int nch = (int)((it.r1 - it.r0 + KC - 1) / KC);
          for (int c = 0; c < nch; ++c) {
            const int64_t r = chunk_start<TN>(it, c, nch, KC);                  // same rotated order as the loader
            const int ksteps = (int)min((int64_t)(KC / 8), (it.r1 - r + 7) / 8);
            mbar_wait((SPLIT == 1 ? RAW : FULL) + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE), sb = sa + A_BYTES;
            for (int j = 0; j < ksteps; ++j) {
              const uint64_t bh = TMA ? make_smem_desc_sw64(sb + j * 32) : make_smem_desc(sb + j * 2 * B_LBO, B_LBO, 128);
              const uint64_t bl = TMA ? make_smem_desc_sw64(sb + LO_OFF + j * 32) : make_smem_desc(sb + LO_OFF + j * 2 * B_LBO, B_LBO, 128);
              for (int hf = 0; hf < halves; ++hf) {
                const uint32_t d = tmem + hf * BN;
                const uint64_t ah = TMA ? make_smem_desc_sw64(sa + hf * A_HALF + j * 32) : make_smem_desc(sa + hf * A_HALF + j * 2 * A_LBO, A_LBO, 128);
                const uint64_t al = TMA ? make_smem_desc_sw64(sa + LO_OFF + hf * A_HALF + j * 32)
                                        : make_smem_desc(sa + LO_OFF + hf * A_HALF + j * 2 * A_LBO, A_LBO, 128);
                if (SPLIT == 3) {
                  umma_tf32(d, al, bh, idesc, (c > 0 || j > 0) ? 1u : 0u);
                  umma_tf32(d, ah, bl, idesc, 1);
                  umma_tf32(d, ah, bh, idesc, 1);
                } else {
                  umma_tf32(d, ah, bh, idesc, (c > 0 || j > 0) ? 1u : 0u);
                }
              }
            }
            umma_commit(EMPTY + 8 * stage);
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
          umma_commit(ACCF);
          umma_commit(ACCF + 8);
          continue;
        }
        const uint32_t buf = iter & 1;
        mbar_wait(ACCE + 8 * buf, ((iter >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + buf * BN;
        const uint32_t idesc = make_idesc_tf32(TM, it.bn, TNT);
        uint32_t acc = 0;
        if (it.r0 >= it.r1) {
          // empty reduction range (TN tail split): nothing to add; still hand the buffer over (epilogue skips it)
        }
        const int nch = it.r0 < it.r1 ? (int)((it.r1 - it.r0 + KC - 1) / KC) : 0;
        for (int c = 0; c < nch; ++c) {
          const int64_t r = chunk_start<TN>(it, c, nch, KC);
          mbar_wait(((TNT && SPLIT == 1) ? RAW : FULL) + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE), sb = sa + A_BYTES;
          const int ksteps = (int)min((int64_t)(KC / 8), (it.r1 - r + 7) / 8);
          for (int j = 0; j < ksteps; ++j) {
            constexpr uint32_t MN_LBO = KC * 128;            // TNT: one 32-column block = KC rows x 128 B
            const uint64_t ah = TNT ? make_smem_desc_mn_sw128(sa + j * 1024, MN_LBO) : make_smem_desc(sa + j * 2 * A_LBO, A_LBO, 128);
            const uint64_t bh = TNT ? make_smem_desc_mn_sw128(sb + j * 1024, MN_LBO) : make_smem_desc(sb + j * 2 * B_LBO, B_LBO, 128);
            if (SPLIT == 3) {
              const uint32_t lo = LO_OFF;
              const uint64_t al = TNT ? make_smem_desc_mn_sw128(sa + lo + j * 1024, MN_LBO) : make_smem_desc(sa + lo + j * 2 * A_LBO, A_LBO, 128);
              const uint64_t bl = TNT ? make_smem_desc_mn_sw128(sb + lo + j * 1024, MN_LBO) : make_smem_desc(sb + lo + j * 2 * B_LBO, B_LBO, 128);
              umma_tf32(d, al, bh, idesc, acc);
              umma_tf32(d, ah, bl, idesc, 1);
              umma_tf32(d, ah, bh, idesc, 1);
            } else {
              umma_tf32(d, ah, bh, idesc, acc);
            }
            acc = 1;
          }
          umma_commit(EMPTY + 8 * stage);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        umma_commit(ACCF + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    // tcgen05.ld gives thread = row, 32 consecutive columns.  NT: transpose each 32x32 chunk through shared memory so that
    // one st.global.v4 of the warp covers 4 rows x 128 contiguous bytes (and the bias / aux / accumulate reads are coalesced
    // the same way).  TN: red.global.add straight from the registers (once per split, not per row tile).
    const int e = warp - EPI_WARP0, q = e & 3, half = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float* const stg = reinterpret_cast<float*>(smem + ST * STAGE + 256) + e * (EPI_ROWS * EPI_LD);
    uint32_t iter = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (p.dact == 0 || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0))) &&
                        (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    const int sub = lane >> 3, c4 = (lane & 7) * 4;   // phase 2: 8 lanes per row, 4 rows per instruction
    for (int64_t w = blockIdx.x; w < p.n_items; w += gridDim.x, ++iter) {
      const Item it = get_item<TN, TMI>(p, w);
     for (int hf = 0; hf < (PAIR ? 2 : 1); ++hf) {
      const uint32_t buf = PAIR ? (uint32_t)hf : (iter & 1);
      mbar_wait(ACCF + 8 * buf, PAIR ? (iter & 1) : ((iter >> 1) & 1));
      tc_fence_after();
      const int ncols = TN ? p.K : p.N;
      const bool has_data = it.r0 < it.r1;
      if (TN) {
        const int64_t row = it.a0 + q * 32 + lane;
        const bool row_ok = row < p.N;
        for (int c0 = half * 32; c0 < it.bn; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_off + buf * BN + c0, v);
          tmem_ld_wait();
          if (!row_ok || !has_data) continue;
          float* crow = p.C + row * p.ldc + it.b0 + c0;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (it.b0 + c0 + c < ncols) atomicAdd(crow + c, __uint_as_float(v[c]));
        }
      } else {
        const int64_t row0 = it.a0 + hf * TM + q * 32;
        for (int c0 = half * 32; (PAIR ? row0 - q * 32 < p.M : true) && c0 < it.bn; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_off + buf * BN + c0, v);
          tmem_ld_wait();
          const int n0 = it.b0 + c0 + c4;
         for (int rh = 0; rh < 32 / EPI_ROWS; ++rh) {          // 16 rows of the chunk per transposition pass
          if ((lane >> 4) == rh) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              *reinterpret_cast<float4*>(stg + (lane & 15) * EPI_LD + c) =
                  make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
          }
          __syncwarp();
          if (vec_ok && n0 + 4 <= ncols) {
            // fast path: 4 independent 16-byte rows per lane; every load of a phase is issued before its first use
            float4 t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = *reinterpret_cast<const float4*>(stg + (i * 4 + sub) * EPI_LD + c4);
            const float4 b = p.bias != nullptr ? ldg4(p.bias + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
            switch (p.act) {
              case 1:
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = bias_act4<1>(t[i], b);
                break;
              case 2:
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = bias_act4<2>(t[i], b);
                break;
              case 3:
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = bias_act4<3>(t[i], b);
                break;
              case 4:
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = bias_act4<4>(t[i], b);
                break;
              default:
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = bias_act4<0>(t[i], b);
            }
            const int64_t rowl = row0 + rh * EPI_ROWS + sub;
            if (p.dact != 0) {
              float4 a[4];
              const float* ap = p.aux + rowl * p.ldaux + n0;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                a[i] = (rowl + i * 4 < p.M) ? *reinterpret_cast<const float4*>(ap + (int64_t)i * 4 * p.ldaux) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                t[i].x *= dact_from_output(a[i].x, p.dact); t[i].y *= dact_from_output(a[i].y, p.dact);
                t[i].z *= dact_from_output(a[i].z, p.dact); t[i].w *= dact_from_output(a[i].w, p.dact);
              }
            }
            float* cp = p.C + rowl * p.ldc + n0;
            if (p.accumulate) {
              float4 o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = (rowl + i * 4 < p.M) ? *reinterpret_cast<const float4*>(cp + (int64_t)i * 4 * p.ldc) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 4; ++i) { t[i].x += o[i].x; t[i].y += o[i].y; t[i].z += o[i].z; t[i].w += o[i].w; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (rowl + i * 4 < p.M) *reinterpret_cast<float4*>(cp + (int64_t)i * 4 * p.ldc) = t[i];
          } else if (n0 < ncols) {
            float bia[4] = {0.f, 0.f, 0.f, 0.f};
            if (p.bias != nullptr) {
#pragma unroll
              for (int el = 0; el < 4; ++el)
                if (n0 + el < ncols) bia[el] = __ldg(p.bias + n0 + el);
            }
            for (int rr = 0; rr < EPI_ROWS; rr += 4) {
              const int rl = rr + sub;
              const int64_t row = row0 + rh * EPI_ROWS + rl;
              if (row >= p.M) continue;
              const float4 t = *reinterpret_cast<const float4*>(stg + rl * EPI_LD + c4);
              const float f[4] = {t.x + bia[0], t.y + bia[1], t.z + bia[2], t.w + bia[3]};
              float* cp = p.C + row * p.ldc + n0;
#pragma unroll
              for (int el = 0; el < 4; ++el) {
                if (n0 + el < ncols) {
                  float x = act_apply(f[el], p.act);
                  if (p.dact != 0) x *= dact_from_output(p.aux[row * p.ldaux + n0 + el], p.dact);
                  cp[el] = p.accumulate ? cp[el] + x : x;
                }
              }
            }
          }
          __syncwarp();
         }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ACCE + 8 * buf);
     }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int SPLIT, bool TN, bool TMA>
constexpr size_t smem_bytes() {
  return (size_t)Cfg<SPLIT, TN, TMA>::ST * Cfg<SPLIT, TN, TMA>::STAGE + 256 + (TN ? 0 : EPI_BYTES);
}


// SM count of the CURRENT device (work-split heuristics only; looked up per call -- an attribute query, no sync)
static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeFn>(ptr);
  }
  return fn;
}
// fp32 [rows, cols] matrix with row stride ld (elements): boxes of box_rows x 16 columns (64 B), SWIZZLE_64B, zero fill out of range
static bool make_tensor_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  TensorMapEncodeFn fn = tensor_map_encoder();
  if (fn == nullptr || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int SPLIT, bool TN, bool TMA>
static int launch_variant(const Params& p, const CUtensorMap& ma, const CUtensorMap& mb, cudaStream_t st) {
  auto kern = gemm_tf32_kernel<SPLIT, TN, TMA>;
  static DeviceOnce once;      // one table per template instantiation
  int num_sms = 0;
  if (int err = device_once(once, "gemm_tf32: device setup", &num_sms, [&] {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<SPLIT, TN, TMA>());
      }))
    return err;
  const int grid = (int)std::min<int64_t>(p.n_items, num_sms);
  kern<<<grid, THREADS, smem_bytes<SPLIT, TN, TMA>(), st>>>(p, ma, mb);
  return check_launch("gemm_tf32_kernel");
}

// fp32 [rows, cols] matrix read MN-major: 3-D view (32 columns, rows, cols / 32) -> boxes of 32 x KC rows x `blocks` column blocks,
// SWIZZLE_128B_ATOM_32B, landing as [block][row][128 B].  cols must be a multiple of 32 (a partial block would read past the row end).
static bool make_tensor_map_mn(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int blocks) {
  TensorMapEncodeFn fn = tensor_map_encoder();
  if (fn == nullptr || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0 || (cols & 31) != 0) return false;
  const cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128u};
  const cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, (cuuint32_t)blocks};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int SPLIT, bool TN>
static int launch(Params p, cudaStream_t st) {
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  // NT work items are TMI rows of A x one B tile; TMI depends on the variant that runs
  auto nt_items = [&](int tmi) { return ((p.M + tmi - 1) / tmi) * p.n_btiles; };
  if (getenv("NSK_GEMM_NO_TMA") == nullptr) {
    if constexpr (!TN) {
      if (make_tensor_map(&ma, p.A, p.M, p.K, p.lda, Cfg<SPLIT, TN, true>::TMI) && make_tensor_map(&mb, p.B, p.N, p.K, p.ldb, BN)) {
        p.n_items = nt_items(Cfg<SPLIT, TN, true>::TMI);
        return launch_variant<SPLIT, TN, true>(p, ma, mb, st);
      }
    } else {
      // TN: A [M, P = p.N] and B [M, Q = p.K]; the reduction index (rows) must fit the int32 TMA coordinate
      if (p.M < (1ll << 31) && make_tensor_map_mn(&ma, p.A, p.M, p.N, p.lda, Cfg<SPLIT, true, true>::KC, TM / 32) &&
          make_tensor_map_mn(&mb, p.B, p.M, p.K, p.ldb, Cfg<SPLIT, true, true>::KC, BN / 32))
        return launch_variant<SPLIT, TN, true>(p, ma, mb, st);
    }
  }
  if constexpr (!TN) p.n_items = nt_items(Cfg<SPLIT, TN, false>::TMI);
  return launch_variant<SPLIT, TN, false>(p, ma, mb, st);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gemm
}  // namespace nsk

extern "C" int nsk_gemm_tf32_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                                const float* bias, int act, const float* aux, int ldaux, int dact, int accumulate, int split,
                                void* stream) {
  using namespace nsk::gemm;
  NSK_REQUIRE(A && B && C, "nsk_gemm_tf32_nt: null pointer");
  NSK_REQUIRE(M >= 0 && N > 0 && K > 0, "nsk_gemm_tf32_nt: bad sizes");
  NSK_REQUIRE((K & 7) == 0, "nsk_gemm_tf32_nt: K must be a multiple of 8 (pad with zeros)");
  NSK_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && lda >= K && ldb >= K && ldc >= N, "nsk_gemm_tf32_nt: leading dimensions");
  NSK_REQUIRE(aligned16(A) && aligned16(B), "nsk_gemm_tf32_nt: A and B must be 16-byte aligned");
  NSK_REQUIRE(act >= 0 && act <= 4 && dact >= 0 && dact <= 4, "nsk_gemm_tf32_nt: act / dact");
  NSK_REQUIRE(dact == 0 || (aux != nullptr && ldaux >= N), "nsk_gemm_tf32_nt: dact needs aux [M, ldaux >= N]");
  NSK_REQUIRE(split == 1 || split == 3, "nsk_gemm_tf32_nt: split must be 1 (tf32) or 3 (3xtf32)");
  if (M == 0) return 0;
  Params p{};
  p.A = A; p.B = B; p.C = C; p.bias = bias;
  p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.act = act; p.accumulate = accumulate;
  p.aux = aux; p.ldaux = ldaux; p.dact = dact;
  p.n_btiles = (N + BN - 1) / BN;
  p.n_atiles = 0;
  p.rows_per_split = 0;
  p.n_items = 0;   // set by launch() once the variant (rows of A per work item) is known
  return split == 3 ? launch<3, false>(p, nsk::as_stream(stream)) : launch<1, false>(p, nsk::as_stream(stream));
}

extern "C" int nsk_gemm_tf32_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int P, int Q,
                                int split, void* stream) {
  using namespace nsk::gemm;
  NSK_REQUIRE(A && B && C, "nsk_gemm_tf32_tn: null pointer");
  NSK_REQUIRE(M >= 0 && P > 0 && Q > 0, "nsk_gemm_tf32_tn: bad sizes");
  NSK_REQUIRE(lda >= P && ldb >= Q && ldc >= Q, "nsk_gemm_tf32_tn: leading dimensions");
  NSK_REQUIRE(split == 1 || split == 3, "nsk_gemm_tf32_tn: split must be 1 (tf32) or 3 (3xtf32)");
  if (M == 0) return 0;
  Params p{};
  p.A = A; p.B = B; p.C = C; p.bias = nullptr;
  p.M = M; p.N = P; p.K = Q; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.act = 0; p.accumulate = 1;
  p.n_atiles = (P + TM - 1) / TM;
  p.n_btiles = (Q + BN - 1) / BN;
  const int per = p.n_atiles * p.n_btiles;
  const int want = std::max(1, sm_count() / per);
  int64_t rows = (M + want - 1) / want;
  rows = std::max<int64_t>(rows, 512);
  rows = (rows + KC_TN - 1) / KC_TN * KC_TN;
  p.rows_per_split = rows;
  p.n_items = ((M + rows - 1) / rows) * per;
  return split == 3 ? launch<3, true>(p, nsk::as_stream(stream)) : launch<1, true>(p, nsk::as_stream(stream));
}

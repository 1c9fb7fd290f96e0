// RENI++ row decode as ONE fused tcgen05 kernel (fp16 operands, fp32 accumulate in TMEM, fp32 LayerNorm epilogues).
//
// Replaces, for frame-sized row batches (the per-ray background radiance of a render, neusky/models/neusky_model.py:535-549, and
// the relighting sweep of BASELINE.json configs[4]), RENIField.get_outputs + Decoder + unnormalise
// (ns_reni/reni/illumination_fields/reni_illumination_field.py:493-573, ns_reni/reni/field_components/transformer_decoder.py:21-155,
// base_spherical_field.py:143-154).  The layer-wise 3xTF32 chain (reni_rows_tc.cu + gemm_tf32.cu) moves every [N,128] fp32
// activation through HBM 4 times per decoder layer (ncu: 13 GEMM + 7 LayerNorm launches at 2.4-5.4 TB/s of DRAM traffic,
// profiles/r02_ncu_hbm_kernels_summary.txt); here a row's activations never leave the SM.
//
// Per row (direction d, latent code k):  pe = NeRF-PE([Z_xy(k) . d_xy (L), d_z, |d_xy|])  (5 (L+2) = 510 columns, L = 100)
//   x  = LN1_0(W_r pe + b_r + a_0(k))                                   a_i(k) = fc_out(value(cond(k))): per-code constants (nsk_reni_prep)
//   for i in 0..5:  h = relu(F0_i x + f0b_i);  y = F2_i h + f2b_i + x;  x = LN2_i(y);  if i < 5: x = LN1_{i+1}(x + a_{i+1}(k))
//   out = exp(W_o x + b_o + log(exp(scale_k)))
//
// One persistent CTA per SM works on PAIRS of 128-row tiles.  Both tiles of a pair run every GEMM against the SAME weight stages
// (weights are fetched from L2 once per 256 rows) and ping-pong between the tensor pipe and the epilogue warps: while tile 0's
// epilogue turns an accumulator into the next A operand, tile 1's MMAs run.  16 GEMM groups of K = 128 per tile
// (4 for the 512-wide first layer, then F0_i / F2_i), 2 x 16 KB weight stages per group, 4-slot bulk-copy ring.
//
//   TMEM (512 columns)   tile t: X = cols [256 t, +128)  first layer / residual stream (F2_i accumulates ONTO x + f2b_i, which the
//                                                         previous epilogue stored there with tcgen05.st);  H = cols [256 t + 128, +128)
//   SMEM   tile t: A_X 32 KB, A_H 32 KB  fp16 A operands (K-major no-swizzle canonical layout); the first layer's four K = 128 quarters
//                  of the positional encoding alternate between them
//          ring 4 x 16 KB weight stages, 20 KB fp32 constants (biases, LayerNorm weights, output head)
//   warp 0  weight producer (one lane)        warp 1  MMA issuer (one lane)        warp 2  TMEM allocator
//   warps 4-7    positional encoding of the NEXT quarter / pair (thread = row; sin / cos on the MUFU after an exact range reduction)
//   warps 8-15   epilogue of tile 0, warps 16-23 epilogue of tile 1: two threads per row (64 columns each, warps e and e + 4 share a TMEM
//                lane quadrant); LayerNorm partial sums cross through shared memory under 64-thread named barriers; registers are moved
//                from the service warps to the epilogue warpgroups with setmaxnreg
#include "reni_common.cuh"
#include "tc_util.cuh"

namespace nsk {
namespace renitc {

using namespace nsk::tc;

constexpr int TM = 128, HID = 128, NLAYER = 6;
constexpr int STAGE_BYTES = 16384;                 // [128 N][64 K] fp16
constexpr int NSLOT = 4;
constexpr int NGROUP = 4 + 2 * NLAYER;             // 16 groups of K = 128 (2 stages each)
constexpr int STAGES_PER_PAIR = 2 * NGROUP;        // 32
constexpr int NUM_THREADS = 768;                   // 8 service / encoding warps + 2 tiles x 8 epilogue warps
constexpr int TILE_BYTES = TM * HID * 2;           // 32768
constexpr int PE_LD = 512;

// constants blob (fp32), also the shared-memory image
constexpr int CB_BR = 0;
constexpr int CB_LAYER0 = 128, CB_LAYER_STRIDE = 6 * 128;   // f0b, f2b, n1w, n1b, n2w, n2b
constexpr int CB_F0B = 0, CB_F2B = 128, CB_N1W = 256, CB_N1B = 384, CB_N2W = 512, CB_N2B = 640;
constexpr int CB_WO = CB_LAYER0 + NLAYER * CB_LAYER_STRIDE;  // [3][128]
constexpr int CB_BO = CB_WO + 3 * 128;                       // [4]
constexpr int CB_FLOATS = CB_BO + 4;                         // 5124
constexpr int64_t WEIGHT_BYTES = (int64_t)STAGES_PER_PAIR * STAGE_BYTES;   // 524288
constexpr int64_t BLOB_BYTES = WEIGHT_BYTES + (int64_t)CB_FLOATS * 4;

constexpr uint32_t OFF_A = 0;                                  // tile t: A_X at t * 65536, A_H at t * 65536 + 32768
constexpr uint32_t OFF_RING = 4 * TILE_BYTES;                  // 131072
constexpr uint32_t OFF_CONST = OFF_RING + NSLOT * STAGE_BYTES; // 196608
constexpr uint32_t OFF_XCHG = OFF_CONST + CB_FLOATS * 4;       // 217104: LayerNorm partial sums [2 tiles][2 halves][128 rows] float4
constexpr uint32_t OFF_BAR = OFF_XCHG + 2 * 2 * 128 * 16;      // 225296
// A_X and A_H of a tile have a READY barrier each: with one shared barrier the positional-encoding warps (quarters 0 and 1 have no
// dependency on the issuer) could complete two phases before the issuer consumed the first, and a parity wait cannot tell phase n from n + 2
enum { B_WFULL = 0, B_WEMPTY = 4, B_AXREADY = 8, B_ACCREADY = 10, B_AXFREE = 12, B_AHFREE = 14, B_ACCFREE = 16, B_AHREADY = 18, B_COUNT = 20 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16;
static_assert(OFF_BAR % 8 == 0, "mbarriers must be 8-byte aligned");
static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory plan exceeds the 227 KB per-CTA limit");

struct Params {
  const float* dirs; const int* row_cam; int64_t N;
  const float* zxy; const float* attn; const float* scale;
  const uint8_t* blob; int L; int log_domain;
  float* out;
  int64_t n_pairs;
  unsigned long long* prof;   // diagnostics: [grid][16] cycle counters (NULL = off; PROF instantiation only)
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
               "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

// 64 fp32 columns of this thread's TMEM lane starting at `taddr`
__device__ __forceinline__ void load_row64(uint32_t taddr, float (&v)[64]) {
  uint32_t u[4][16];
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld16(taddr + c * 16, u[c]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c * 16 + j] = __uint_as_float(u[c][j]);
}

// v += a: 64 per-column constants through 16-byte broadcast loads (shared memory) / read-only global loads
__device__ __forceinline__ void add64(float (&v)[64], const float* __restrict__ a) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 x = a4[j];
    v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
  }
}
__device__ __forceinline__ void add64_ldg(float (&v)[64], const float* __restrict__ a) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 x = __ldg(a4 + j);
    v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
  }
}

// LayerNorm over a row of 128 held by TWO threads (64 columns each, same TMEM lane, warps w and w + 4 of the tile's epilogue group):
// biased variance, eps 1e-5 (torch.nn.LayerNorm), two-pass like the reference; the partial sums cross through shared memory
// (`mine` / `theirs`: this thread's and its partner's float4 slot) under a 64-thread named barrier.  w / b: this half's 64 columns.
__device__ __forceinline__ void layernorm_half(float (&v)[64], const float* __restrict__ w, const float* __restrict__ b, float4* mine, const float4* theirs, int bar_id) {
  // one exchange per LayerNorm: (sum, sum of squares) of this half; var = E[v^2] - mean^2 in fp32 (the rows here are O(1) with
  // |mean| < std, so the cancellation costs ~1e-7 relative; the two-pass form needed a second named-barrier round trip)
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 64; ++j) { s[j & 3] += v[j]; q[j & 3] = fmaf(v[j], v[j], q[j & 3]); }
  const float s_mine = (s[0] + s[1]) + (s[2] + s[3]), q_mine = (q[0] + q[1]) + (q[2] + q[3]);
  mine->x = s_mine; mine->y = q_mine;
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  const float mean = (s_mine + theirs->x) * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, (q_mine + theirs->y) * (1.0f / 128.0f)), 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float shift = -mean * rstd;
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // the partner has read this slot before the next LayerNorm rewrites it
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 ww = w4[j], bb = b4[j];
    v[4 * j + 0] = fmaf(fmaf(v[4 * j + 0], rstd, shift), ww.x, bb.x);
    v[4 * j + 1] = fmaf(fmaf(v[4 * j + 1], rstd, shift), ww.y, bb.y);
    v[4 * j + 2] = fmaf(fmaf(v[4 * j + 2], rstd, shift), ww.z, bb.z);
    v[4 * j + 3] = fmaf(fmaf(v[4 * j + 3], rstd, shift), ww.w, bb.w);
  }
}

// this thread's 64 columns of a row: fp16 image into the next A operand + the fp32 residual base (x + bias) back into TMEM
__device__ __forceinline__ void store_half(const float (&v)[64], uint8_t* a_dst /* tile base + first chunk of this half + row * 16 */, uint32_t tmem_x,
                                           const float* __restrict__ resid_bias) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(a_dst + c * (TM * 16)) = make_uint4(pack_h2(v[c * 8], v[c * 8 + 1]), pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                                                  pack_h2(v[c * 8 + 4], v[c * 8 + 5]), pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
  const float4* r4 = reinterpret_cast<const float4*>(resid_bias);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t u[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 rb = r4[c * 4 + j];
      u[4 * j + 0] = __float_as_uint(v[c * 16 + 4 * j + 0] + rb.x);
      u[4 * j + 1] = __float_as_uint(v[c * 16 + 4 * j + 1] + rb.y);
      u[4 * j + 2] = __float_as_uint(v[c * 16 + 4 * j + 2] + rb.z);
      u[4 * j + 3] = __float_as_uint(v[c * 16 + 4 * j + 3] + rb.w);
    }
    tmem_st16(tmem_x + c * 16, u);
  }
  tmem_st_wait();
}

template <bool PROF>
__global__ void __launch_bounds__(NUM_THREADS, 1) reni_rows_fused_kernel(const Params P) {
#define RCLK() (PROF ? clock64() : 0ll)
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  const float* cb = reinterpret_cast<const float*>(smem + OFF_CONST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pair0 = blockIdx.x, pair_step = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bars + 8 * (B_WFULL + i), 1); mbar_init(bars + 8 * (B_WEMPTY + i), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bars + 8 * (B_AXREADY + t), 8);     // the tile's 8 epilogue warps arrive once each, the 4 encoding warps twice each
      mbar_init(bars + 8 * (B_AHREADY + t), 8);
      mbar_init(bars + 8 * (B_ACCREADY + t), 1);    // tcgen05.commit
      mbar_init(bars + 8 * (B_AXFREE + t), 1);
      mbar_init(bars + 8 * (B_AHFREE + t), 1);
      mbar_init(bars + 8 * (B_ACCFREE + t), 8);
    }
    fence_barrier_init();
  }
  {  // constants -> shared memory (once per CTA)
    const float* src = reinterpret_cast<const float*>(P.blob + WEIGHT_BYTES);
    float* dst = reinterpret_cast<float*>(smem + OFF_CONST);
    for (int i = threadIdx.x; i < CB_FLOATS; i += NUM_THREADS) dst[i] = __ldg(src + i);
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0 && lane == 0) {
      // ================================ weight producer ================================
      const uint64_t pol = l2_policy_evict_last();
      uint32_t G = 0;
      for (int64_t pr = pair0; pr < P.n_pairs; pr += pair_step) {
#pragma unroll 1
        for (int s = 0; s < STAGES_PER_PAIR; ++s, ++G) {
          const uint32_t slot = G & (NSLOT - 1), ph = (G / NSLOT) & 1u;
          mbar_wait(bars + 8 * (B_WEMPTY + slot), ph ^ 1u);
          mbar_arrive_expect_tx(bars + 8 * (B_WFULL + slot), STAGE_BYTES);
          bulk_g2s_hint(sbase + OFF_RING + slot * STAGE_BYTES, P.blob + (int64_t)s * STAGE_BYTES, STAGE_BYTES, bars + 8 * (B_WFULL + slot), pol);
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      const uint32_t idesc = make_idesc_f16(TM, HID);
      const uint64_t desc_hi = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)((TM * 16) >> 4) << 16);   // SBO 128 B, LBO = 128 rows * 16 B
      uint32_t G = 0, ph_axready[2] = {0, 0}, ph_ahready[2] = {0, 0}, ph_accfree[2] = {0, 0};
      long long t_a = 0, t_w = 0, t_f = 0, t_e1 = 0, t_e2 = 0;
      const long long t_begin = RCLK();
      for (int64_t pr = pair0; pr < P.n_pairs; pr += pair_step) {
#pragma unroll 1
        for (int g = 0; g < NGROUP; ++g, G += 2) {
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            long long c0 = RCLK();
            if (g == 0) { mbar_wait(bars + 8 * (B_ACCFREE + t), ph_accfree[t] ^ 1u); ph_accfree[t] ^= 1u; }   // previous pair's last epilogue has read X
            long long c1 = RCLK();
            const bool first_layer = g < 4;
            const bool use_ah = first_layer ? (g & 1) : (((g - 4) & 1) != 0);                 // quarters alternate A_X / A_H; F0 reads A_X, F2 reads A_H
            if (use_ah) { mbar_wait(bars + 8 * (B_AHREADY + t), ph_ahready[t]); ph_ahready[t] ^= 1u; }
            else { mbar_wait(bars + 8 * (B_AXREADY + t), ph_axready[t]); ph_axready[t] ^= 1u; }
            long long c2 = RCLK();
            if (t == 0) {
#pragma unroll
              for (int s = 0; s < 2; ++s) mbar_wait(bars + 8 * (B_WFULL + ((G + s) & (NSLOT - 1))), ((G + s) / NSLOT) & 1u);
            }
            if (PROF) {
              const long long c3 = RCLK();
              t_f += c1 - c0; t_w += c3 - c2;
              if (g < 4) t_a += c2 - c1;                   // waiting for the positional encoding
              else if ((g - 4) & 1) t_e1 += c2 - c1;       // F2_i waits for the relu epilogue
              else t_e2 += c2 - c1;                        // F0_i waits for the LayerNorm epilogue
            }
            tc_fence_after();
            const bool to_h = !first_layer && (((g - 4) & 1) == 0);                           // F0_i -> H ; first layer and F2_i -> X
            const uint32_t a_base = sbase + OFF_A + (uint32_t)t * (2 * TILE_BYTES) + (use_ah ? TILE_BYTES : 0);
            const uint32_t d_tmem = tmem + (uint32_t)t * 256 + (to_h ? 128 : 0);
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                const uint32_t slot = (G + s) & (NSLOT - 1);
                uint64_t ad = desc_hi | (uint64_t)(((a_base + s * (8 * TM * 16)) >> 4) & 0x3FFF);
                uint64_t bd = desc_hi | (uint64_t)(((sbase + OFF_RING + slot * STAGE_BYTES) >> 4) & 0x3FFF);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t acc = (first_layer ? (g > 0) : !to_h) || s > 0 || j > 0;     // F2_i accumulates onto x + f2b_i; first-layer quarters chain
                  umma_ss(d_tmem, ad, bd, idesc, acc ? 1u : 0u);
                  ad += (2 * TM * 16) >> 4;       // one K = 16 step = 2 chunks of 8 columns
                  bd += (2 * HID * 16) >> 4;
                }
                if (t == 1) umma_commit(bars + 8 * (B_WEMPTY + slot));
              }
              if (g == 0) umma_commit(bars + 8 * (B_AXFREE + t));
              else if (g == 1) umma_commit(bars + 8 * (B_AHFREE + t));
              else if (g >= 3) umma_commit(bars + 8 * (B_ACCREADY + t));
              if (g == NGROUP - 2) umma_commit(bars + 8 * (B_AXFREE + t));
              if (g == NGROUP - 1) umma_commit(bars + 8 * (B_AHFREE + t));
            }
            __syncwarp();
          }
        }
      }
      if (PROF && P.prof && lane == 0) {
        unsigned long long* o = P.prof + blockIdx.x * 16;
        o[0] = RCLK() - t_begin; o[1] = t_a; o[2] = t_e1; o[3] = t_e2; o[4] = t_w; o[5] = t_f;
      }
    }
  } else if (warp < 8) {
    // ================================ positional encoding ================================
    // K-column layout of the first layer for L = 100 (the contraction is order-free; packing.pack_reni_fused permutes W_r the same
    // way).  Quarter q (128 columns = one GEMM group) is SELF-CONTAINED: it holds the 25 latent inputs j = 25 q .. 25 q + 24
    //   cols 4 jj + {0,1,2,3} = sin a, sin 4a, cos a, cos 4a  (a = 2 pi xin_j; the 4a pair by two double-angle steps)
    //   cols 100 + jj         = xin_j
    //   cols 125 .. 127       = entries 3 q .. 3 q + 2 of [sin, sin4, cos, cos4, x](d_z) ++ [sin, sin4, cos, cos4, x](|d_xy|) ++ [0, 0]
    // so a thread loads the quarter's 25 (z_x, z_y) pairs up front (25 independent 8-byte loads in flight), and one input costs two
    // MUFU ops for its five columns.
    const int row = (warp - 4) * 32 + lane;
    uint32_t ph_ax[2] = {0, 0}, ph_ah[2] = {0, 0};
    long long t_pw = 0;
    const long long t_pbegin = RCLK();
    auto trig = [](float x, float& s1, float& s4, float& c1, float& c4) {
      const float u = x - rintf(x);                          // sin(2 pi x): exact reduction to [-1/2, 1/2], then the MUFU
      const float a = 6.283185307179586f * u;
      s1 = __sinf(a); c1 = __cosf(a);
      const float s2 = 2.f * s1 * c1, c2 = fmaf(-2.f * s1, s1, 1.f);
      s4 = 2.f * s2 * c2; c4 = fmaf(-2.f * s2, s2, 1.f);
    };
    for (int64_t pr = pair0; pr < P.n_pairs; pr += pair_step) {
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t grow = min(pr * 256 + t * 128 + row, P.N - 1);
          const int k = P.row_cam ? __ldg(P.row_cam + grow) : 0;
          const float dx = __ldg(P.dirs + grow * 3), dy = __ldg(P.dirs + grow * 3 + 1), dz = __ldg(P.dirs + grow * 3 + 2);
          const float2* zp = reinterpret_cast<const float2*>(P.zxy) + (int64_t)k * 100 + 25 * q;
          float xj[25];
#pragma unroll
          for (int jj = 0; jj < 25; ++jj) { const float2 z = __ldg(zp + jj); xj[jj] = fmaf(z.x, dx, z.y * dy); }
          // this quarter's three columns of the two extra inputs d_z and |d_xy| (reni_illumination_field.py:219-246)
          float ex[3];
          {
            const float dxy = sqrtf(dx * dx + dy * dy);
            float zs1, zs4, zc1, zc4, ys1, ys4, yc1, yc4;
            trig(dz, zs1, zs4, zc1, zc4);
            trig(dxy, ys1, ys4, yc1, yc4);
            ex[0] = q == 0 ? zs1 : q == 1 ? zc4 : q == 2 ? ys4 : dxy;
            ex[1] = q == 0 ? zs4 : q == 1 ? dz : q == 2 ? yc1 : 0.f;
            ex[2] = q == 0 ? zc1 : q == 1 ? ys1 : q == 2 ? yc4 : 0.f;
          }
          // the buffer this quarter goes to must have been read by the MMAs that used it last
          const long long pc0 = RCLK();
          if (q & 1) { mbar_wait_backoff(bars + 8 * (B_AHFREE + t), ph_ah[t] ^ 1u, 256); ph_ah[t] ^= 1u; }
          else { mbar_wait_backoff(bars + 8 * (B_AXFREE + t), ph_ax[t] ^ 1u, 256); ph_ax[t] ^= 1u; }
          if (PROF) t_pw += RCLK() - pc0;
          uint8_t* dst = smem + OFF_A + t * (2 * TILE_BYTES) + ((q & 1) ? TILE_BYTES : 0) + row * 16;
#pragma unroll
          for (int c = 0; c < 12; ++c) {
            float a0, a1, a2, a3, b0, b1, b2, b3;
            trig(xj[2 * c], a0, a1, a2, a3);
            trig(xj[2 * c + 1], b0, b1, b2, b3);
            *reinterpret_cast<uint4*>(dst + c * (TM * 16)) = make_uint4(pack_h2(a0, a1), pack_h2(a2, a3), pack_h2(b0, b1), pack_h2(b2, b3));
          }
          {
            float a0, a1, a2, a3;
            trig(xj[24], a0, a1, a2, a3);
            *reinterpret_cast<uint4*>(dst + 12 * (TM * 16)) = make_uint4(pack_h2(a0, a1), pack_h2(a2, a3), pack_h2(xj[0], xj[1]), pack_h2(xj[2], xj[3]));
          }
          *reinterpret_cast<uint4*>(dst + 13 * (TM * 16)) = make_uint4(pack_h2(xj[4], xj[5]), pack_h2(xj[6], xj[7]), pack_h2(xj[8], xj[9]), pack_h2(xj[10], xj[11]));
          *reinterpret_cast<uint4*>(dst + 14 * (TM * 16)) = make_uint4(pack_h2(xj[12], xj[13]), pack_h2(xj[14], xj[15]), pack_h2(xj[16], xj[17]), pack_h2(xj[18], xj[19]));
          *reinterpret_cast<uint4*>(dst + 15 * (TM * 16)) = make_uint4(pack_h2(xj[20], xj[21]), pack_h2(xj[22], xj[23]), pack_h2(xj[24], ex[0]), pack_h2(ex[1], ex[2]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { mbar_arrive(bars + 8 * (((q & 1) ? B_AHREADY : B_AXREADY) + t)); mbar_arrive(bars + 8 * (((q & 1) ? B_AHREADY : B_AXREADY) + t)); }
        }
      }
    }
    if (PROF && P.prof && row == 0) { P.prof[blockIdx.x * 16 + 6] = RCLK() - t_pbegin; P.prof[blockIdx.x * 16 + 7] = t_pw; }
  } else {
    // ================================ epilogue (warps 8-15: tile 0, warps 16-23: tile 1) ================================
    // Two threads per row: warps e and e + 4 of a tile's group own the same TMEM lanes (lane quadrant = warp % 4) and split the 128
    // columns.  With one thread per row (128 live registers, one epilogue warp per SM sub-partition and tile) the epilogue ran at an IPC
    // of ~0.2 and was 2.5x longer than the MMAs it feeds (profiles/r02_reni_fused_phase_cycles.log).
    reg_inc<88>();
    const int e = warp - 8;
    const int t = e >> 3, half = (e >> 2) & 1;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_x = tmem + lane_off + (uint32_t)t * 256 + half * 64, tm_h = tm_x + 128;
    uint8_t* ax = smem + OFF_A + t * (2 * TILE_BYTES) + (half * 8) * (TM * 16) + row * 16;      // this half's first 8-column chunk
    uint8_t* ah = ax + TILE_BYTES;
    float4* xc = reinterpret_cast<float4*>(smem + OFF_XCHG);
    float4* mine = xc + (t * 2 + half) * 128 + row;
    const float4* theirs = xc + (t * 2 + (half ^ 1)) * 128 + row;
    const int bar_id = 1 + t * 4 + (warp & 3);                                                  // the two warps sharing these 32 rows
    const int co = half * 64;                                                                   // first column of this half
    const uint32_t bar_acc = bars + 8 * (B_ACCREADY + t), bar_ax = bars + 8 * (B_AXREADY + t), bar_ah = bars + 8 * (B_AHREADY + t);
    uint32_t ph_acc = 0;
    long long t_ew = 0;
    const long long t_ebegin = RCLK();
    auto wait_acc = [&]() {
      const long long c0 = RCLK();
      mbar_wait_backoff(bar_acc, ph_acc); ph_acc ^= 1u;
      if (PROF) t_ew += RCLK() - c0;
      tc_fence_after();
    };
    auto publish = [&](uint32_t bar) {      // A operand (and TMEM residual) written -> the issuer may run the next GEMM of this tile
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    for (int64_t pr = pair0; pr < P.n_pairs; pr += pair_step) {
      const int64_t grow_raw = pr * 256 + t * 128 + row;
      const int64_t grow = min(grow_raw, P.N - 1);
      const int k = P.row_cam ? __ldg(P.row_cam + grow) : 0;
      const float* at = P.attn + (int64_t)k * (NLAYER * HID) + co;
      float v[64];
      // ---- first layer: x = LN1_0(W_r pe + b_r + a_0) ----
      wait_acc();
      load_row64(tm_x, v);
      add64(v, cb + CB_BR + co);
      add64_ldg(v, at);
      layernorm_half(v, cb + CB_LAYER0 + CB_N1W + co, cb + CB_LAYER0 + CB_N1B + co, mine, theirs, bar_id);
      store_half(v, ax, tm_x, cb + CB_LAYER0 + CB_F2B + co);
      publish(bar_ax);
#pragma unroll 1
      for (int i = 0; i < NLAYER; ++i) {
        const float* cl = cb + CB_LAYER0 + i * CB_LAYER_STRIDE + co;
        // ---- h = relu(F0_i x + f0b_i) ----
        wait_acc();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t u[16];
          tmem_ld16(tm_h + c * 16, u);
          tmem_ld_wait();
          uint32_t pk[8];
          const float4* fb = reinterpret_cast<const float4*>(cl + CB_F0B + c * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b = fb[j];
            pk[2 * j] = pack_h2(fmaxf(__uint_as_float(u[4 * j]) + b.x, 0.f), fmaxf(__uint_as_float(u[4 * j + 1]) + b.y, 0.f));
            pk[2 * j + 1] = pack_h2(fmaxf(__uint_as_float(u[4 * j + 2]) + b.z, 0.f), fmaxf(__uint_as_float(u[4 * j + 3]) + b.w, 0.f));
          }
          *reinterpret_cast<uint4*>(ah + (2 * c) * (TM * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(ah + (2 * c + 1) * (TM * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        publish(bar_ah);
        // ---- y = F2_i h + (x + f2b_i) (accumulated in TMEM);  x = LN2_i(y) ----
        wait_acc();
        load_row64(tm_x, v);
        if (i == NLAYER - 1) {        // X of this tile is free for the next pair's first layer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + 8 * (B_ACCFREE + t));
        }
        layernorm_half(v, cl + CB_N2W, cl + CB_N2B, mine, theirs, bar_id);
        if (i + 1 < NLAYER) {
          const float* cn = cl + CB_LAYER_STRIDE;
          add64_ldg(v, at + (i + 1) * HID);
          layernorm_half(v, cn + CB_N1W, cn + CB_N1B, mine, theirs, bar_id);
          store_half(v, ax, tm_x, cn + CB_F2B);
          publish(bar_ax);
        }
      }
      // ---- head: 128 -> 3, + scale in the log domain, exp (reni_illumination_field.py:561-565, base_spherical_field.py:143-154) ----
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        const float4* wo = reinterpret_cast<const float4*>(cb + CB_WO + c * 128 + co);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 w = wo[j];
          s4[0] = fmaf(v[4 * j], w.x, s4[0]); s4[1] = fmaf(v[4 * j + 1], w.y, s4[1]); s4[2] = fmaf(v[4 * j + 2], w.z, s4[2]); s4[3] = fmaf(v[4 * j + 3], w.w, s4[3]);
        }
        o[c] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      if (half == 1) { mine->x = o[0]; mine->y = o[1]; mine->z = o[2]; }
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] += cb[CB_BO + c];
        o[0] += theirs->x; o[1] += theirs->y; o[2] += theirs->z;
        if (P.scale) {
          const float sc = expf(__ldg(P.scale + k));
#pragma unroll
          for (int c = 0; c < 3; ++c) o[c] = P.log_domain ? (o[c] + logf(sc)) : (o[c] * sc);
        }
        if (grow_raw < P.N) {
#pragma unroll
          for (int c = 0; c < 3; ++c) P.out[grow_raw * 3 + c] = P.log_domain == 1 ? expf(o[c]) : o[c];
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // the partner has read the head partials before the next pair reuses the slot
    }
    if (PROF && P.prof && row == 0 && half == 0) { P.prof[blockIdx.x * 16 + 8 + 2 * t] = RCLK() - t_ebegin; P.prof[blockIdx.x * 16 + 9 + 2 * t] = t_ew; }
  }
#undef RCLK

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace renitc
}  // namespace nsk

extern "C" int64_t nsk_reni_fused_weights_bytes(void) { return nsk::renitc::BLOB_BYTES; }

// Diagnostics only (not part of the public ABI): per-CTA cycle counters [grid][16] (scripts/reni_fused_phase_profile.py names the slots).
static unsigned long long* g_nsk_reni_fused_prof = nullptr;
extern "C" void nsk_debug_set_reni_fused_prof(unsigned long long* buf) { g_nsk_reni_fused_prof = buf; }

extern "C" int nsk_reni_rows_fused_fwd(const float* dirs, const int* row_cam, int64_t N, const float* zxy, const float* attn, const float* scale,
                                       const void* fused_weights, int latent_dim, int log_domain, float* out, void* stream) {
  using namespace nsk::renitc;
  if (N == 0) return 0;
  NSK_REQUIRE(dirs && zxy && attn && fused_weights && out, "nsk_reni_rows_fused_fwd: null pointer");
  NSK_REQUIRE(latent_dim == 100, "nsk_reni_rows_fused_fwd: the fused kernel is specialised for latent_dim = 100 (the RENI++ decoder NeuSky ships)");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(fused_weights) & 15) == 0, "nsk_reni_rows_fused_fwd: weight blob must be 16-byte aligned");
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "nsk_reni_rows_fused_fwd: device setup", &num_sms, [] {
        cudaError_t e = cudaFuncSetAttribute(reni_rows_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(reni_rows_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        return e;
      }))
    return err;
  Params P;
  P.dirs = dirs; P.row_cam = row_cam; P.N = N; P.zxy = zxy; P.attn = attn; P.scale = scale;
  P.blob = reinterpret_cast<const uint8_t*>(fused_weights); P.L = latent_dim; P.log_domain = log_domain; P.out = out;
  P.n_pairs = (N + 255) / 256;
  P.prof = g_nsk_reni_fused_prof;
  const int64_t grid = P.n_pairs < num_sms ? P.n_pairs : num_sms;
  if (P.prof) reni_rows_fused_kernel<true><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  else reni_rows_fused_kernel<false><<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, nsk::as_stream(stream)>>>(P);
  return nsk::check_launch("reni_rows_fused_kernel");
}

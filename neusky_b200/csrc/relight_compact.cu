// Compact relighting cache and the streaming pass over it.
//
// The relighting sweep of BASELINE.json configs[4] re-shades fixed geometry under new RENI++ latent codes (reference:
// neusky/models/neusky_model.py:1896-1980, publication/render_animation.py:188-221).  With geometry fixed, everything a new illumination
// needs from the render is, per ray, the visibility-weighted Lambert coefficients  H[r, j, c] = vis(r, j) * sum_s w_s albedo_s,c
// clamp(n_s . l_j) / count_s  (nsk_lambert_collapse_sel x visibility): rgb_lin[r, c] = sum_j H[r, j, c] L[j, c].
// Round 1 kept H as fp32 [R, D, 3] for every ray (7.1 GB per 1280x720 frame at D = 642).  Here:
//   * only rays that hit something (accumulation > 0) own a row: `rows` [Rs] int32 -> ray index (sky rays shade to 0);
//   * a row is fp16, channel-planar [3][DP] with DP = D rounded up to 16, normalised by its own maximum (`hscale` [Rs] fp32), so the
//     fp16 mantissa is spent on the row's dynamic range: 3.9 KB per hit ray;
//   * the pass is a skinny matrix product per channel, [rows x DP] . [DP x NL], and it must stay HBM-bound: a SIMT version (fp16 -> fp32
//     conversions + FFMA, 27 instructions per direction and row pair for eight codes) was instruction-issue bound at 2.2 TB/s.  So each
//     warp takes 16 rows and feeds them to warp-level mma (m16n8k16, fp16 x fp16 -> fp32): the eight bytes a lane loads per row and
//     16-direction block ARE its A fragment (inside a block the directions are stored in fragment order: lane t of a quad holds
//     directions 2t, 2t+1, 2t+8, 2t+9), the radiance tables sit in shared memory in B-fragment order, and one read of the cache serves
//     up to 32 latent codes (four n = 8 blocks).  ~7 instructions per 2048 multiply-adds; tcgen05 would add nothing here (the tensor
//     work is 1 % of what the pipe can do at HBM speed), so the kernel has no TMEM / mbarrier machinery.
// HBM-bound: 6 DP + 4 bytes per hit ray and pass of up to 32 latent codes.
#include "nsk_common.cuh"

namespace nsk {

constexpr int RC_WARPS = 16;      // warps per block of the pass kernel (one block per SM: the tables take up to 158 KB of shared memory)
constexpr int RC_NL = 32;         // latent codes per pass
constexpr int RC_TS = 40;         // words per (k pair) row of the table: 32 codes + 8 pad -> the four k pairs of a quad land in distinct banks
constexpr int RC_PACK_WARPS = 8;

// position p of a 16-direction block -> direction offset inside the block (A-fragment order of mma.m16n8k16, see above)
__device__ __forceinline__ int rc_frag_dir(int p) {
  const int t = p >> 2, j = p & 3;
  return 2 * t + (j & 1) + 8 * (j >> 1);
}

// one warp per cache row: row max -> scale, fp16 planar image in fragment order
__global__ void __launch_bounds__(RC_PACK_WARPS * 32)
relight_pack_h16_kernel(const float* __restrict__ H, const int32_t* __restrict__ rows, int64_t Rs, int D, int DP, __half* __restrict__ H16,
                        float* __restrict__ hscale) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RC_PACK_WARPS + (threadIdx.x >> 5);
  if (i >= Rs) return;
  const float* hr = H + (int64_t)rows[i] * D * 3;
  const int n = D * 3;
  float m = 0.f;
  for (int e = lane; e < n; e += 32) m = fmaxf(m, fabsf(hr[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float inv = m > 0.f ? 1.0f / m : 0.f;
  if (lane == 0) hscale[i] = m;
  __half* dst = H16 + i * (int64_t)(3 * DP);
  for (int e = lane; e < 3 * DP; e += 32) {
    const int c = e / DP, pos = e - c * DP;
    const int d = (pos & ~15) + rc_frag_dir(pos & 15);
    dst[e] = __float2half_rn(d < D ? hr[d * 3 + c] * inv : 0.f);
  }
}

__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rgb_lin[l, rows[i], c] = hscale[i] * sum_d H16[i, c, d] * L[l, d, c]   for l < NL (<= 32) illuminations of this pass.
// NB = number of n = 8 code blocks actually computed (1..4).
template <int NB>
__global__ void __launch_bounds__(RC_WARPS * 32, 1)
relight_h16_kernel(const uint2* __restrict__ H16, const float* __restrict__ hscale, const int32_t* __restrict__ rows, int64_t Rs, int64_t R, int D,
                   int DP, const float* __restrict__ radiance /* [NL, D, 3] */, int NL, float* __restrict__ rgb_lin /* [NL, R, 3] */) {
  extern __shared__ __align__(16) uint8_t rc_smem[];
  // radiance as fp16 scaled by 1 / max (HDR tables span several decades), in B-fragment order:
  // word [(c * steps + s) * 8 + kp][code] = (L[code, 16 s + 2 kp, c], L[code, 16 s + 2 kp + 1, c]), row stride RC_TS words
  uint32_t* tab = reinterpret_cast<uint32_t*>(rc_smem);
  __shared__ float s_max[RC_NL];
  if (threadIdx.x < RC_NL) s_max[threadIdx.x] = 0.f;
  __syncthreads();
  for (int l = 0; l < NL; ++l) {
    float m = 0.f;
    for (int e = threadIdx.x; e < D * 3; e += blockDim.x) m = fmaxf(m, fabsf(__ldg(radiance + (size_t)l * D * 3 + e)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(&s_max[l]), __float_as_int(m));     // non-negative floats order like ints
  }
  __syncthreads();
  const int steps = DP >> 4;
  for (int e = threadIdx.x; e < 3 * steps * 8 * (NB * 8); e += blockDim.x) {
    const int code = e % (NB * 8);
    const int kp = (e / (NB * 8)) & 7;
    const int cs = e / (NB * 8 * 8);           // c * steps + s
    const int c = cs / steps, s = cs - c * steps;
    const int d0 = 16 * s + 2 * kp;
    const float mx = code < NL ? s_max[code] : 0.f;
    const float inv = mx > 0.f ? 1.0f / mx : 0.f;
    const float v0 = (code < NL && d0 < D) ? __ldg(radiance + ((size_t)code * D + d0) * 3 + c) * inv : 0.f;
    const float v1 = (code < NL && d0 + 1 < D) ? __ldg(radiance + ((size_t)code * D + d0 + 1) * 3 + c) * inv : 0.f;
    const __half2 h = __floats2half2_rn(v0, v1);
    tab[(size_t)(cs * 8 + kp) * RC_TS + code] = *reinterpret_cast<const uint32_t*>(&h);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t ntiles = (Rs + 15) >> 4;
  const int64_t warps_total = (int64_t)gridDim.x * RC_WARPS;
  const int row_u2 = (3 * DP) >> 2;                  // uint2 (4 halfs) per cache row
  for (int64_t tile = (int64_t)blockIdx.x * RC_WARPS + (threadIdx.x >> 5); tile < ntiles; tile += warps_total) {
    const int64_t r0 = tile * 16 + g, r1 = r0 + 8;
    const uint2* p0 = H16 + min(r0, Rs - 1) * row_u2 + t;       // lane t of a quad: halfs 4t..4t+3 of every 16-direction block
    const uint2* p1 = H16 + min(r1, Rs - 1) * row_u2 + t;
    float acc[3][NB][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][nb][j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint2* q0 = p0 + c * (DP >> 2);
      const uint2* q1 = p1 + c * (DP >> 2);
      const uint32_t* tb = tab + (size_t)(c * steps) * 8 * RC_TS + t * RC_TS + g;
      constexpr int U = 8;                            // 16 independent 8-byte loads in flight per lane
      for (int s0 = 0; s0 < steps; s0 += U) {
        uint2 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          va[u] = vb[u] = make_uint2(0u, 0u);
          if (s0 + u < steps) {
            va[u] = __ldcs(q0 + (s0 + u) * 4);        // streamed once per pass
            vb[u] = __ldcs(q1 + (s0 + u) * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (s0 + u < steps) {
            const uint32_t a[4] = {va[u].x, vb[u].x, va[u].y, vb[u].y};     // (row g, k 2t..), (row g+8, k 2t..), (row g, k 2t+8..), (row g+8, k 2t+8..)
            const uint32_t* ts = tb + (size_t)(s0 + u) * 8 * RC_TS;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) mma_m16n8k16(acc[c][nb], a, ts[nb * 8], ts[4 * RC_TS + nb * 8]);
          }
        }
      }
    }
    // D fragment: acc[.][nb][0,1] = (row g, codes 8 nb + 2t, + 1), acc[.][nb][2,3] = (row g + 8, same codes)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t i = half ? r1 : r0;
      if (i < Rs) {
        const float hs = hscale[i];
        const int64_t ray = rows[i];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int l = nb * 8 + 2 * t + j;
            if (l < NL) {
              const float sc = hs * s_max[l];
              float* o = rgb_lin + ((size_t)l * R + ray) * 3;
              o[0] = acc[0][nb][half * 2 + j] * sc; o[1] = acc[1][nb][half * 2 + j] * sc; o[2] = acc[2][nb][half * 2 + j] * sc;
            }
          }
      }
    }
  }
}

}  // namespace nsk

extern "C" int nsk_relight_pack_h16(const float* H, const int32_t* rows, int64_t Rs, int D, void* H16, float* hscale, void* stream) {
  if (Rs == 0) return 0;
  NSK_REQUIRE(D >= 1 && H && rows && H16 && hscale, "nsk_relight_pack_h16: null pointer / D");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(H16) & 15) == 0, "nsk_relight_pack_h16: H16 must be 16-byte aligned");
  const int DP = (D + 15) & ~15;
  const int64_t blocks = (Rs + nsk::RC_PACK_WARPS - 1) / nsk::RC_PACK_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_relight_pack_h16: too many rows for one launch");
  nsk::relight_pack_h16_kernel<<<(unsigned)blocks, nsk::RC_PACK_WARPS * 32, 0, nsk::as_stream(stream)>>>(H, rows, Rs, D, DP, reinterpret_cast<__half*>(H16), hscale);
  return nsk::check_launch("relight_pack_h16_kernel");
}

namespace nsk {
template <int NB>
static int relight_h16_launch(const void* H16, const float* hscale, const int32_t* rows, int64_t Rs, int64_t R, int D, int DP, const float* radiance, int NL,
                              float* rgb_lin, cudaStream_t st) {
  const size_t smem = (size_t)3 * (DP >> 4) * 8 * RC_TS * sizeof(uint32_t);
  static DeviceOnce once;
  int num_sms = 0;
  if (int err = device_once(once, "nsk_relight_h16_multi: shared memory opt-in", &num_sms, [] {
        return cudaFuncSetAttribute(relight_h16_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      }))
    return err;
  NSK_REQUIRE(smem <= 200 * 1024, "nsk_relight_h16_multi: too many directions for the shared-memory radiance tables");
  const int64_t want = ((Rs + 15) / 16 + RC_WARPS - 1) / RC_WARPS;
  const unsigned grid = (unsigned)(want < num_sms ? want : num_sms);      // persistent: the radiance tables are staged once per block
  relight_h16_kernel<NB><<<grid, RC_WARPS * 32, smem, st>>>(reinterpret_cast<const uint2*>(H16), hscale, rows, Rs, R, D, DP, radiance, NL, rgb_lin);
  return check_launch("relight_h16_kernel");
}
}  // namespace nsk

extern "C" int nsk_relight_h16_multi(const void* H16, const float* hscale, const int32_t* rows, int64_t Rs, int64_t R, int D, const float* radiance,
                                     int NL, float* rgb_lin, void* stream) {
  if (R == 0 || NL == 0) return 0;
  NSK_REQUIRE(D >= 1 && NL >= 1 && radiance && rgb_lin && (Rs == 0 || (H16 && hscale && rows)), "nsk_relight_h16_multi: null pointer / sizes");
  cudaStream_t st = nsk::as_stream(stream);
  // rays without a cache row (sky) shade to zero
  if (cudaMemsetAsync(rgb_lin, 0, (size_t)NL * R * 3 * sizeof(float), st) != cudaSuccess) return nsk::fail("nsk_relight_h16_multi", "memset");
  if (Rs == 0) return 0;
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(H16) & 15) == 0, "nsk_relight_h16_multi: H16 must be 16-byte aligned");
  const int DP = (D + 15) & ~15;
  const size_t n = (size_t)D * 3;
  for (int l = 0; l < NL; l += nsk::RC_NL) {
    const int nl = NL - l < nsk::RC_NL ? NL - l : nsk::RC_NL;
    const float* rad = radiance + (size_t)l * n;
    float* out = rgb_lin + (size_t)l * R * 3;
    int err;
    if (nl <= 8) err = nsk::relight_h16_launch<1>(H16, hscale, rows, Rs, R, D, DP, rad, nl, out, st);
    else if (nl <= 16) err = nsk::relight_h16_launch<2>(H16, hscale, rows, Rs, R, D, DP, rad, nl, out, st);
    else err = nsk::relight_h16_launch<4>(H16, hscale, rows, Rs, R, D, DP, rad, nl, out, st);
    if (err) return err;
  }
  return 0;
}

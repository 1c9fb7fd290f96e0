// Compact relighting cache and the streaming pass over it.
//
// The relighting sweep of BASELINE.json configs[4] re-shades fixed geometry under new RENI++ latent codes (reference:
// neusky/models/neusky_model.py:1896-1980, publication/render_animation.py:188-221).  With geometry fixed, everything a new illumination
// needs from the render is, per ray, the visibility-weighted Lambert coefficients  H[r, j, c] = vis(r, j) * sum_s w_s albedo_s,c
// clamp(n_s . l_j) / count_s  (nsk_lambert_collapse_sel x visibility): rgb_lin[r, c] = sum_j H[r, j, c] L[j, c].
// Round 1 kept H as fp32 [R, D, 3] for every ray (7.1 GB per 1280x720 frame at D = 642).  Here:
//   * only rays that hit something (accumulation > 0) own a row: `rows` [Rs] int32 -> ray index (sky rays shade to 0);
//   * a row is fp16, channel-planar [3][DP] with DP = D rounded up to 8 (16-byte aligned rows and channel planes), normalised by
//     its own maximum (`hscale` [Rs] fp32), so the fp16 mantissa is spent on the row's dynamic range: 3.9 KB per hit ray;
//   * one pass streams the cache ONCE for four illuminations (radiance tables staged in shared memory as fp16, interleaved per
//     direction so one 16-byte load feeds four FMAs x two directions).
// HBM-bound: 6 DP + 4 bytes per hit ray and pass of four latent codes.
#include "nsk_common.cuh"

namespace nsk {

constexpr int RC_WARPS = 8;
constexpr int RC_NL = 4;

// one warp per cache row: row max -> scale, fp16 planar image
__global__ void __launch_bounds__(RC_WARPS * 32)
relight_pack_h16_kernel(const float* __restrict__ H, const int32_t* __restrict__ rows, int64_t Rs, int D, int DP, __half* __restrict__ H16,
                        float* __restrict__ hscale) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RC_WARPS + (threadIdx.x >> 5);
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
    const int c = e / DP, d = e - c * DP;
    dst[e] = __float2half_rn(d < D ? hr[d * 3 + c] * inv : 0.f);
  }
}

// rgb_lin[l, rows[i], c] = hscale[i] * sum_d H16[i, c, d] * L[l, d, c]   for l < NL (<= 4) illuminations of this pass
__global__ void __launch_bounds__(RC_WARPS * 32)
relight_h16_kernel(const uint4* __restrict__ H16, const float* __restrict__ hscale, const int32_t* __restrict__ rows, int64_t Rs, int64_t R, int D,
                   int DP, const float* __restrict__ radiance /* [NL, D, 3] */, int NL, float* __restrict__ rgb_lin /* [NL, R, 3] */) {
  extern __shared__ __align__(16) uint8_t rc_smem[];
  // radiance as fp16 scaled by 1 / max (HDR tables span several decades), [3][DP][4 illuminations]: 8 bytes per (channel, direction)
  __half* rs = reinterpret_cast<__half*>(rc_smem);
  __shared__ float s_max[RC_NL];
  if (threadIdx.x < RC_NL) {
    s_max[threadIdx.x] = 0.f;
  }
  __syncthreads();
  {
    float m[RC_NL] = {0.f, 0.f, 0.f, 0.f};
    for (int e = threadIdx.x; e < D * 3; e += blockDim.x)
#pragma unroll
      for (int l = 0; l < RC_NL; ++l)
        if (l < NL) m[l] = fmaxf(m[l], fabsf(__ldg(radiance + (size_t)l * D * 3 + e)));
#pragma unroll
    for (int l = 0; l < RC_NL; ++l) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m[l] = fmaxf(m[l], __shfl_xor_sync(0xffffffffu, m[l], o));
      if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(&s_max[l]), __float_as_int(m[l]));     // non-negative floats order like ints
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * DP; e += blockDim.x) {
    const int c = e / DP, d = e - c * DP;
#pragma unroll
    for (int l = 0; l < RC_NL; ++l) {
      const float mx = s_max[l];
      const float v = (l < NL && d < D && mx > 0.f) ? __ldg(radiance + ((size_t)l * D + d) * 3 + c) / mx : 0.f;
      rs[(size_t)e * RC_NL + l] = __float2half_rn(v);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int chunks = DP >> 3;                        // 16-byte chunks (8 directions) per channel plane
  const int64_t warps_total = (int64_t)gridDim.x * RC_WARPS;
  const uint4* rs4 = reinterpret_cast<const uint4*>(rs);
  for (int64_t i = (int64_t)blockIdx.x * RC_WARPS + (threadIdx.x >> 5); i < Rs; i += warps_total) {
    const uint4* hr = H16 + i * (int64_t)(3 * chunks);
    float acc[3][RC_NL];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int l = 0; l < RC_NL; ++l) acc[c][l] = 0.f;
      for (int k = lane; k < chunks; k += 32) {
        const uint4 hv = __ldcs(hr + c * chunks + k);                   // streamed once per pass
        const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
        const uint4* rp = rs4 + ((size_t)c * DP + (size_t)k * 8) / 2;   // 8 directions x 4 illuminations x 2 B = 64 B = 4 x 16 B
#pragma unroll
        for (int p = 0; p < 4; ++p) {                                   // direction pair p: directions 2p, 2p + 1 of the chunk
          const float2 h = __half22float2(h2[p]);
          const uint4 rv = rp[p];                                       // [dir 2p: l0 l1 l2 l3 | dir 2p+1: l0 l1 l2 l3]
          const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
          const float2 a01 = __half22float2(r2[0]), a23 = __half22float2(r2[1]), b01 = __half22float2(r2[2]), b23 = __half22float2(r2[3]);
          acc[c][0] = fmaf(h.x, a01.x, acc[c][0]); acc[c][1] = fmaf(h.x, a01.y, acc[c][1]);
          acc[c][2] = fmaf(h.x, a23.x, acc[c][2]); acc[c][3] = fmaf(h.x, a23.y, acc[c][3]);
          acc[c][0] = fmaf(h.y, b01.x, acc[c][0]); acc[c][1] = fmaf(h.y, b01.y, acc[c][1]);
          acc[c][2] = fmaf(h.y, b23.x, acc[c][2]); acc[c][3] = fmaf(h.y, b23.y, acc[c][3]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int l = 0; l < RC_NL; ++l) acc[c][l] = warp_sum(acc[c][l]);
    if (lane == 0) {
      const float hs = hscale[i];
      const int64_t ray = rows[i];
      for (int l = 0; l < NL; ++l) {
        float* o = rgb_lin + ((size_t)l * R + ray) * 3;
        const float s = hs * s_max[l];
        o[0] = acc[0][l] * s; o[1] = acc[1][l] * s; o[2] = acc[2][l] * s;
      }
    }
  }
}

}  // namespace nsk

extern "C" int nsk_relight_pack_h16(const float* H, const int32_t* rows, int64_t Rs, int D, void* H16, float* hscale, void* stream) {
  if (Rs == 0) return 0;
  NSK_REQUIRE(D >= 1 && H && rows && H16 && hscale, "nsk_relight_pack_h16: null pointer / D");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(H16) & 15) == 0, "nsk_relight_pack_h16: H16 must be 16-byte aligned");
  const int DP = (D + 7) & ~7;
  const int64_t blocks = (Rs + nsk::RC_WARPS - 1) / nsk::RC_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_relight_pack_h16: too many rows for one launch");
  nsk::relight_pack_h16_kernel<<<(unsigned)blocks, nsk::RC_WARPS * 32, 0, nsk::as_stream(stream)>>>(H, rows, Rs, D, DP, reinterpret_cast<__half*>(H16), hscale);
  return nsk::check_launch("relight_pack_h16_kernel");
}

extern "C" int nsk_relight_h16_multi(const void* H16, const float* hscale, const int32_t* rows, int64_t Rs, int64_t R, int D, const float* radiance,
                                     int NL, float* rgb_lin, void* stream) {
  if (R == 0 || NL == 0) return 0;
  NSK_REQUIRE(D >= 1 && NL >= 1 && radiance && rgb_lin && (Rs == 0 || (H16 && hscale && rows)), "nsk_relight_h16_multi: null pointer / sizes");
  cudaStream_t st = nsk::as_stream(stream);
  // rays without a cache row (sky) shade to zero
  if (cudaMemsetAsync(rgb_lin, 0, (size_t)NL * R * 3 * sizeof(float), st) != cudaSuccess) return nsk::fail("nsk_relight_h16_multi", "memset");
  if (Rs == 0) return 0;
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(H16) & 15) == 0, "nsk_relight_h16_multi: H16 must be 16-byte aligned");
  const int DP = (D + 7) & ~7;
  const size_t smem = (size_t)3 * DP * nsk::RC_NL * sizeof(__half);
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "nsk_relight_h16_multi: shared memory opt-in", &num_sms, [] {
        return cudaFuncSetAttribute(nsk::relight_h16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      }))
    return err;
  NSK_REQUIRE(smem <= 96 * 1024, "nsk_relight_h16_multi: too many directions for the shared-memory radiance tables");
  const int64_t want = (Rs + nsk::RC_WARPS - 1) / nsk::RC_WARPS;
  const int64_t cap = (int64_t)num_sms * 4;                       // persistent: the radiance tables are staged once per block
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const size_t n = (size_t)D * 3;
  for (int l = 0; l < NL; l += nsk::RC_NL) {
    const int nl = NL - l < nsk::RC_NL ? NL - l : nsk::RC_NL;
    nsk::relight_h16_kernel<<<grid, nsk::RC_WARPS * 32, smem, st>>>(reinterpret_cast<const uint4*>(H16), hscale, rows, Rs, R, D, DP, radiance + (size_t)l * n, nl,
                                                                    rgb_lin + (size_t)l * R * 3);
  }
  return nsk::check_launch("relight_h16_kernel");
}

// Compact relighting cache and the streaming pass over it.
//
// The relighting sweep of BASELINE.json configs[4] re-shades fixed geometry under new RENI++ latent codes (reference:
// neusky/models/neusky_model.py:1896-1980, publication/render_animation.py:188-221).  With geometry fixed, everything a new illumination
// needs from the render is, per ray, the visibility-weighted Lambert coefficients  H[r, j, c] = vis(r, j) * sum_s w_s albedo_s,c
// clamp(n_s . l_j) / count_s  (nsk_lambert_collapse_sel x visibility): rgb_lin[r, c] = sum_j H[r, j, c] L[j, c].
// Round 1 kept H as fp32 [R, D, 3] for every ray (7.1 GB per 1280x720 frame at D = 642).  Here:
//   * only rays that hit something (accumulation > 0) own a row: `rows` [Rs] int32 -> ray index (sky rays shade to 0);
//   * a row is fp16, channel-planar [3][DP] with DP = D rounded up to 8 (16-byte aligned rows and channel planes), normalised by
//     its own maximum (`hscale` [Rs] fp32), so the fp16 mantissa is spent on the row's dynamic range: 3.9 KB per hit ray.
//     Inside a channel plane the directions are STRIDED over the 16-byte chunks: element e of chunk k is direction k + (DP/8) e,
//     so the lanes of a warp (lane = chunk) touch consecutive directions of the shared-memory radiance table at every step
//     (conflict-free 16-byte shared loads; with direction-contiguous chunks the lanes were 64 B apart, a 4-way bank conflict);
//   * one pass streams the cache ONCE for EIGHT illuminations: the radiance tables sit in shared memory as fp16 [channel][direction][8]
//     (one 16-byte load = one direction of all eight tables), and a warp works on TWO rows at a time so every table load feeds 16 FMAs.
//     The first version (four codes per pass, one row per warp) ran at 1.4 TB/s of cache bytes: its table reads cost 4x the
//     shared-memory bandwidth of the global bytes they served, 4-way conflicted.
// HBM-bound: 6 DP + 4 bytes per hit ray and pass of eight latent codes.
#include "nsk_common.cuh"

namespace nsk {

constexpr int RC_WARPS = 8;
constexpr int RC_NL = 8;
constexpr int RC_KI = 3;      // 16-byte chunks per lane and load batch (3 x 32 chunks = 768 directions per channel in one batch)

// one warp per cache row: row max -> scale, fp16 planar image with strided direction order (see above)
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
  const int chunks = DP >> 3;
  for (int e = lane; e < 3 * DP; e += 32) {
    const int c = e / DP, pos = e - c * DP;
    const int d = (pos >> 3) + chunks * (pos & 7);          // chunk k = pos / 8, element pos % 8 -> direction k + chunks * element
    dst[e] = __float2half_rn(d < D ? hr[d * 3 + c] * inv : 0.f);
  }
}

// rgb_lin[l, rows[i], c] = hscale[i] * sum_d H16[i, c, d] * L[l, d, c]   for l < NL (<= 8) illuminations of this pass
__global__ void __launch_bounds__(RC_WARPS * 32)
relight_h16_kernel(const uint4* __restrict__ H16, const float* __restrict__ hscale, const int32_t* __restrict__ rows, int64_t Rs, int64_t R, int D,
                   int DP, const float* __restrict__ radiance /* [NL, D, 3] */, int NL, float* __restrict__ rgb_lin /* [NL, R, 3] */) {
  extern __shared__ __align__(16) uint8_t rc_smem[];
  // radiance as fp16 scaled by 1 / max (HDR tables span several decades), [3][DP][8 illuminations]: 16 bytes per (channel, direction)
  __half* rs = reinterpret_cast<__half*>(rc_smem);
  __shared__ float s_max[RC_NL];
  if (threadIdx.x < RC_NL) s_max[threadIdx.x] = 0.f;
  __syncthreads();
  {
    float m[RC_NL];
#pragma unroll
    for (int l = 0; l < RC_NL; ++l) m[l] = 0.f;
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
  const int chunks = DP >> 3;                        // 16-byte chunks per channel plane
  const int64_t warps_total = (int64_t)gridDim.x * RC_WARPS;
  const uint4* rs4 = reinterpret_cast<const uint4*>(rs);
  const int64_t npairs = (Rs + 1) >> 1;
  for (int64_t pi = (int64_t)blockIdx.x * RC_WARPS + (threadIdx.x >> 5); pi < npairs; pi += warps_total) {
    const int64_t i0 = 2 * pi, i1 = min(i0 + 1, Rs - 1);
    const uint4* hr0 = H16 + i0 * (int64_t)(3 * chunks);
    const uint4* hr1 = H16 + i1 * (int64_t)(3 * chunks);
    // accumulators [row 2][illumination 8][channel 3], channel fastest: the transposed reduction below leaves lane pair g with the
    // three channels of (row g / 8, illumination g % 8)
    float acc[48];
#pragma unroll
    for (int j = 0; j < 48; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      for (int kb = 0; kb < chunks; kb += 32 * RC_KI) {
        // every global load of this (channel, block of 96 chunks) is in flight before the first FMA: 3 KB per warp
        uint4 hv0[RC_KI], hv1[RC_KI];
#pragma unroll
        for (int u = 0; u < RC_KI; ++u) {
          const int k = kb + lane + 32 * u;
          hv0[u] = hv1[u] = make_uint4(0u, 0u, 0u, 0u);
          if (k < chunks) {
            hv0[u] = __ldcs(hr0 + c * chunks + k);                       // streamed once per pass
            hv1[u] = __ldcs(hr1 + c * chunks + k);
          }
        }
#pragma unroll
        for (int u = 0; u < RC_KI; ++u) {
          const int k = kb + lane + 32 * u;
          if (k >= chunks) continue;
          const __half2* a2 = reinterpret_cast<const __half2*>(&hv0[u]);
          const __half2* b2 = reinterpret_cast<const __half2*>(&hv1[u]);
          const uint4* rp = rs4 + (size_t)c * DP + k;                    // direction k + chunks * e -> stride `chunks` entries per element
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 ha = __half22float2(a2[e2]), hb = __half22float2(b2[e2]);
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint4 rv = rp[(size_t)(2 * e2 + w) * chunks];
              const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
              const float h0 = w ? ha.y : ha.x, h1 = w ? hb.y : hb.x;
#pragma unroll
              for (int l2 = 0; l2 < 4; ++l2) {
                const float2 rr = __half22float2(r2[l2]);
                acc[(2 * l2) * 3 + c] = fmaf(h0, rr.x, acc[(2 * l2) * 3 + c]);
                acc[(2 * l2 + 1) * 3 + c] = fmaf(h0, rr.y, acc[(2 * l2 + 1) * 3 + c]);
                acc[24 + (2 * l2) * 3 + c] = fmaf(h1, rr.x, acc[24 + (2 * l2) * 3 + c]);
                acc[24 + (2 * l2 + 1) * 3 + c] = fmaf(h1, rr.y, acc[24 + (2 * l2 + 1) * 3 + c]);
              }
            }
          }
        }
      }
    }
    // transposed warp reduction: every step halves the values a lane carries (48 -> 24 -> 12 -> 6 -> 3), 48 shuffles instead of 240
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int m = 16 >> step, half = 24 >> step;
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float send = up ? acc[j] : acc[j + half];
        const float keep = up ? acc[j + half] : acc[j];
        acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    if ((lane & 1) == 0) {
      const int g = lane >> 1, rsel = g >> 3, l = g & 7;
      const int64_t i = rsel ? i0 + 1 : i0;
      if (l < NL && i < Rs) {
        const float sc = hscale[i] * s_max[l];
        float* o = rgb_lin + ((size_t)l * R + rows[i]) * 3;
        o[0] = acc[0] * sc; o[1] = acc[1] * sc; o[2] = acc[2] * sc;
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
  const int64_t want = ((Rs + 1) / 2 + nsk::RC_WARPS - 1) / nsk::RC_WARPS;
  const int64_t cap = (int64_t)num_sms * 2;                       // persistent, two resident blocks per SM (122 registers): the radiance tables are staged once per block
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const size_t n = (size_t)D * 3;
  for (int l = 0; l < NL; l += nsk::RC_NL) {
    const int nl = NL - l < nsk::RC_NL ? NL - l : nsk::RC_NL;
    nsk::relight_h16_kernel<<<grid, nsk::RC_WARPS * 32, smem, st>>>(reinterpret_cast<const uint4*>(H16), hscale, rows, Rs, R, D, DP, radiance + (size_t)l * n, nl,
                                                                    rgb_lin + (size_t)l * R * 3);
  }
  return nsk::check_launch("relight_h16_kernel");
}

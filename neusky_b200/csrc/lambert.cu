// Lambertian pre-pass (cosine normaliser + un-occluded directions) and final blend + sRGB.
// Replaces the parts of RGBLambertianRendererWithVisibility.render_and_combine_rgb that do not
// depend on the DDF (neusky/model_components/renderers.py:93-106, 122-128, 173-174) and
// linear_to_sRGB (neusky/utils/utils.py:11-31).  The reference feeds these einsums with
// [R*S, D, 3] tensors of repeated directions / colours (neusky_model.py:512-525); here the [D,3]
// direction set and the [K,D,3] radiance table are read as they are.
#include "nsk_common.cuh"

namespace nsk {

constexpr int LP_WARPS = 8;

// one warp per sample: lanes stride over the D directions
__global__ void __launch_bounds__(LP_WARPS * 32)
lambert_prep_kernel(const float* __restrict__ normals, const float* __restrict__ wa, int64_t R, int S,
                    const float* __restrict__ dirs, const uint8_t* __restrict__ ddf_mask, int D,
                    const float* __restrict__ radiance, const int32_t* __restrict__ cam, float unocc_vis,
                    float* __restrict__ inv_count, float* __restrict__ rgb_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* rad = radiance + (int64_t)(cam ? cam[ray] : 0) * D * 3;
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  for (int s = 0; s < S; ++s) {
    const int64_t i = ray * S + s;
    const float nx = normals[i * 3], ny = normals[i * 3 + 1], nz = normals[i * 3 + 2];
    float cnt = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int j = lane; j < D; j += 32) {
      float c = nx * dirs[j * 3] + ny * dirs[j * 3 + 1] + nz * dirs[j * 3 + 2];
      c = fminf(fmaxf(c, 0.f), 1.f);                 // renderers.py:98
      cnt += (c > 0.f) ? 1.f : 0.f;                  // renderers.py:101
      if (!ddf_mask[j]) { a0 += c * rad[j * 3]; a1 += c * rad[j * 3 + 1]; a2 += c * rad[j * 3 + 2]; }
    }
    cnt = warp_sum(cnt); a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    const float ic = 1.0f / (cnt > 0.f ? cnt : 1.0f);  // renderers.py:104-106
    if (lane == 0) inv_count[i] = ic;
    const float k = ic * unocc_vis;
    o0 += wa[i * 3] * a0 * k; o1 += wa[i * 3 + 1] * a1 * k; o2 += wa[i * 3 + 2] * a2 * k;
  }
  if (lane == 0) { rgb_lin[ray * 3] = o0; rgb_lin[ray * 3 + 1] = o1; rgb_lin[ray * 3 + 2] = o2; }
}

// Relighting pass: the Lambertian sum of lambert_prep + K4 with the per-ray visibility taken from a cache instead of
// the DDF (fixed geometry, new illumination: neusky/models/neusky_model.py:1896-1980 re-renders everything per frame).
// one warp per ray; vis_sel [R, Dp] holds the visibility of the directions with ddf_mask == 1, in mask order.
__global__ void __launch_bounds__(LP_WARPS * 32)
lambert_relight_kernel(const float* __restrict__ normals, const float* __restrict__ wa, const float* __restrict__ inv_count,
                       int64_t R, int S, const float* __restrict__ dirs, const int32_t* __restrict__ sel_index, int D, int Dp,
                       const float* __restrict__ radiance, const int32_t* __restrict__ cam, const float* __restrict__ vis_sel,
                       float unocc_vis, float* __restrict__ rgb_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* rad = radiance + (int64_t)(cam ? cam[ray] : 0) * D * 3;
  const float* vr = vis_sel + ray * Dp;
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  for (int j = lane; j < D; j += 32) {
    const float lx = dirs[j * 3], ly = dirs[j * 3 + 1], lz = dirs[j * 3 + 2];
    const int sj = sel_index[j];                       // position in the masked set, or -1
    const float v = sj >= 0 ? vr[sj] : unocc_vis;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int s = 0; s < S; ++s) {
      const int64_t i = ray * S + s;
      float c = normals[i * 3] * lx + normals[i * 3 + 1] * ly + normals[i * 3 + 2] * lz;
      c = fminf(fmaxf(c, 0.f), 1.f) * inv_count[i];
      c0 = fmaf(wa[i * 3], c, c0); c1 = fmaf(wa[i * 3 + 1], c, c1); c2 = fmaf(wa[i * 3 + 2], c, c2);
    }
    o0 = fmaf(c0 * v, rad[j * 3], o0); o1 = fmaf(c1 * v, rad[j * 3 + 1], o1); o2 = fmaf(c2 * v, rad[j * 3 + 2], o2);
  }
  o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2);
  if (lane == 0) { rgb_lin[ray * 3] = o0; rgb_lin[ray * 3 + 1] = o1; rgb_lin[ray * 3 + 2] = o2; }
}

__device__ __forceinline__ float srgb(float c) {
  // neusky/utils/utils.py:25-30
  const float v = (c <= 0.0031308f) ? 12.92f * c : 1.055f * powf(fabsf(c), 1.0f / 2.4f) - 0.055f;
  return fminf(fmaxf(v, 0.f), 1.f);
}

__global__ void shade_finalize_kernel(const float* __restrict__ rgb_lin, const float* __restrict__ bg,
                                      const float* __restrict__ acc, int64_t R, float* __restrict__ rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 3) return;
  const int64_t r = i / 3;
  rgb[i] = srgb(rgb_lin[i] + bg[i] * (1.0f - acc[r]));  // renderers.py:127-128 (+ eval clamp, a no-op after srgb)
}

}  // namespace nsk

extern "C" int nsk_lambert_prep(const float* normals, const float* wa, int64_t R, int S, const float* dirs,
                                const uint8_t* ddf_mask, int D, const float* radiance, const int32_t* cam,
                                float unoccluded_vis, float* inv_count, float* rgb_lin, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1 && D >= 1, "nsk_lambert_prep: S and D must be >= 1");
  NSK_REQUIRE(normals && wa && dirs && ddf_mask && radiance && inv_count && rgb_lin, "nsk_lambert_prep: null pointer");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_lambert_prep: too many rays for one launch");
  nsk::lambert_prep_kernel<<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      normals, wa, R, S, dirs, ddf_mask, D, radiance, cam, unoccluded_vis, inv_count, rgb_lin);
  return nsk::check_launch("lambert_prep_kernel");
}

extern "C" int nsk_shade_finalize(const float* rgb_lin, const float* bg, const float* acc, int64_t R, int training,
                                  float* rgb, void* stream) {
  (void)training;  // linear_to_sRGB already clamps to [0,1]; the eval-only clamp is then the identity
  if (R == 0) return 0;
  NSK_REQUIRE(rgb_lin && bg && acc && rgb, "nsk_shade_finalize: null pointer");
  const int64_t n = R * 3;
  nsk::shade_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(rgb_lin, bg, acc, R, rgb);
  return nsk::check_launch("shade_finalize_kernel");
}

extern "C" int nsk_lambert_relight(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                                   const float* dirs, const int32_t* sel_index, int D, int Dp, const float* radiance,
                                   const int32_t* cam, const float* vis_sel, float unoccluded_vis, float* rgb_lin,
                                   void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1 && D >= 1, "nsk_lambert_relight: S and D must be >= 1");
  NSK_REQUIRE(normals && wa && inv_count && dirs && sel_index && radiance && rgb_lin && (vis_sel || Dp == 0), "nsk_lambert_relight: null pointer");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_lambert_relight: too many rays for one launch");
  nsk::lambert_relight_kernel<<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      normals, wa, inv_count, R, S, dirs, sel_index, D, Dp, radiance, cam, vis_sel, unoccluded_vis, rgb_lin);
  return nsk::check_launch("lambert_relight_kernel");
}

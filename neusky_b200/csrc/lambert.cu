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

// Same outputs with one THREAD per sample (full renders: S = 48..128 samples per ray, one camera per launch): the direction
// set sits in shared memory as {l, ddf_mask} and the radiance of the un-masked directions as a second float4 array, the
// direction loop is uniform across the block (broadcast LDS, no shuffles, no per-sample warp reduction) and the per-ray sum
// over samples is one warp reduction + one red.global.add per warp.  ~2.5x fewer issued instructions than the warp-per-ray
// form above, which stays for S < 32 (the shading microbench has S = 1) and for mixed-camera batches.
constexpr int LPS_THREADS = 256;

__global__ void __launch_bounds__(LPS_THREADS)
lambert_prep_samples_kernel(const float* __restrict__ normals, const float* __restrict__ wa, int64_t N, int S,
                            const float* __restrict__ dirs, const uint8_t* __restrict__ ddf_mask, int D,
                            const float* __restrict__ radiance, float unocc_vis, float* __restrict__ inv_count,
                            float* __restrict__ rgb_lin) {
  extern __shared__ float4 s_lp[];
  float4* sd = s_lp;            // [D] {lx, ly, lz, masked ? 1 : 0}
  float4* sr = s_lp + D;        // [D] {Lr, Lg, Lb, 0}
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    sd[j] = make_float4(dirs[j * 3], dirs[j * 3 + 1], dirs[j * 3 + 2], ddf_mask[j] ? 1.f : 0.f);
    sr[j] = make_float4(radiance[j * 3], radiance[j * 3 + 1], radiance[j * 3 + 2], 0.f);
  }
  __syncthreads();
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < N;
  const int64_t i = live ? i0 : N - 1;
  const float nx = normals[i * 3], ny = normals[i * 3 + 1], nz = normals[i * 3 + 2];
  float cnt = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
  for (int j = 0; j < D; ++j) {
    const float4 l = sd[j];
    float c = nx * l.x + ny * l.y + nz * l.z;
    c = fminf(fmaxf(c, 0.f), 1.f);                   // renderers.py:98
    cnt += (c > 0.f) ? 1.f : 0.f;                    // renderers.py:101
    if (l.w == 0.f) {                                // block-uniform: directions that do not go through the DDF
      const float4 L = sr[j];
      a0 = fmaf(c, L.x, a0); a1 = fmaf(c, L.y, a1); a2 = fmaf(c, L.z, a2);
    }
  }
  const float ic = 1.0f / (cnt > 0.f ? cnt : 1.0f);  // renderers.py:104-106
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  if (live) {
    inv_count[i] = ic;
    const float k = ic * unocc_vis;
    o0 = wa[i * 3] * a0 * k; o1 = wa[i * 3 + 1] * a1 * k; o2 = wa[i * 3 + 2] * a2 * k;
  }
  const int64_t ray = i / S;
  const int64_t ray0 = __shfl_sync(0xffffffffu, ray, 0);
  if (__all_sync(0xffffffffu, ray == ray0)) {
    o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2);
    if ((threadIdx.x & 31) == 0) { atomicAdd(rgb_lin + ray * 3, o0); atomicAdd(rgb_lin + ray * 3 + 1, o1); atomicAdd(rgb_lin + ray * 3 + 2, o2); }
  } else if (live) {
    atomicAdd(rgb_lin + ray * 3, o0); atomicAdd(rgb_lin + ray * 3 + 1, o1); atomicAdd(rgb_lin + ray * 3 + 2, o2);
  }
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

// Collapsed relighting cache.  Everything in the Lambertian sum except the light colours is fixed once the
// geometry is: H[r, j, c] = vis[r, j] * sum_s wa[r, s, c] * clamp01(n[r, s] . l_j) * inv_count[r, s], so a new illumination costs
// one pass over H: rgb_lin[r, c] = sum_j H[r, j, c] * L[j, c] -- D x 3 floats streamed per ray, no per-sample data, no DDF.
// one warp per ray, lanes over directions (same arithmetic order as lambert_relight_kernel up to the final product)
__global__ void __launch_bounds__(LP_WARPS * 32)
lambert_collapse_kernel(const float* __restrict__ normals, const float* __restrict__ wa, const float* __restrict__ inv_count,
                        int64_t R, int S, const float* __restrict__ dirs, const int32_t* __restrict__ sel_index, int D, int Dp,
                        const float* __restrict__ vis_sel, float unocc_vis, float* __restrict__ H) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* vr = vis_sel + ray * Dp;
  float* hr = H + ray * D * 3;
  for (int j = lane; j < D; j += 32) {
    const float lx = dirs[j * 3], ly = dirs[j * 3 + 1], lz = dirs[j * 3 + 2];
    const int sj = sel_index[j];
    const float v = sj >= 0 ? vr[sj] : unocc_vis;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int s = 0; s < S; ++s) {
      const int64_t i = ray * S + s;
      float c = normals[i * 3] * lx + normals[i * 3 + 1] * ly + normals[i * 3 + 2] * lz;
      c = fminf(fmaxf(c, 0.f), 1.f) * inv_count[i];
      c0 = fmaf(wa[i * 3], c, c0); c1 = fmaf(wa[i * 3 + 1], c, c1); c2 = fmaf(wa[i * 3 + 2], c, c2);
    }
    hr[j * 3] = c0 * v; hr[j * 3 + 1] = c1 * v; hr[j * 3 + 2] = c2 * v;
  }
}

// G[r, j, c] = sum_s wa[r, s, c] * clamp01(n[r, s] . l_j) * inv_count[r, s] for the Dp directions that go through the DDF: the
// per-pair Lambert coefficient K4 otherwise recomputes from the S samples in its tail (nsk_sky_shade_tc2_fwd with S = 0 reads
// this table instead).  One block per ray: the ray's samples are staged in shared memory as {n, .} and {wa * inv_count, .},
// each thread owns up to 3 directions in registers and walks the samples with broadcast loads.
constexpr int LCS_THREADS = 128;
constexpr int LCS_DPT = 3;
__global__ void __launch_bounds__(LCS_THREADS)
lambert_collapse_sel_kernel(const float* __restrict__ normals, const float* __restrict__ wa, const float* __restrict__ inv_count,
                            int64_t R, int S, const float* __restrict__ dirs_sel, int Dp, float* __restrict__ G) {
  extern __shared__ float4 s_lc[];
  float4* sn = s_lc;          // [S] {nx, ny, nz, 0}
  float4* sw = s_lc + S;      // [S] {w0, w1, w2, 0} * inv_count
  const int64_t ray = blockIdx.x;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int64_t i = ray * S + s;
    const float ic = inv_count[i];
    sn[s] = make_float4(normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2], 0.f);
    sw[s] = make_float4(wa[i * 3] * ic, wa[i * 3 + 1] * ic, wa[i * 3 + 2] * ic, 0.f);
  }
  __syncthreads();
  for (int j0 = threadIdx.x; j0 < Dp; j0 += LCS_THREADS * LCS_DPT) {
    float lx[LCS_DPT], ly[LCS_DPT], lz[LCS_DPT], a[LCS_DPT][3];
#pragma unroll
    for (int d = 0; d < LCS_DPT; ++d) {
      const int j = min(j0 + d * LCS_THREADS, Dp - 1);
      lx[d] = dirs_sel[j * 3]; ly[d] = dirs_sel[j * 3 + 1]; lz[d] = dirs_sel[j * 3 + 2];
      a[d][0] = a[d][1] = a[d][2] = 0.f;
    }
#pragma unroll 2
    for (int s = 0; s < S; ++s) {
      const float4 n = sn[s], w = sw[s];
#pragma unroll
      for (int d = 0; d < LCS_DPT; ++d) {
        float c = n.x * lx[d] + n.y * ly[d] + n.z * lz[d];
        c = fminf(fmaxf(c, 0.f), 1.f);
        a[d][0] = fmaf(w.x, c, a[d][0]); a[d][1] = fmaf(w.y, c, a[d][1]); a[d][2] = fmaf(w.z, c, a[d][2]);
      }
    }
#pragma unroll
    for (int d = 0; d < LCS_DPT; ++d) {
      const int j = j0 + d * LCS_THREADS;
      if (j < Dp) {
        float* g = G + (ray * Dp + j) * 3;
        g[0] = a[d][0]; g[1] = a[d][1]; g[2] = a[d][2];
      }
    }
  }
}

// rgb_lin[r, c] = sum_j H[r, j, c] * L[cam(r)][j, c]: one warp per ray streaming its D*3 floats (HBM-bound: 12 D bytes per ray)
__global__ void __launch_bounds__(LP_WARPS * 32)
relight_collapsed_kernel(const float* __restrict__ H, int64_t R, int D, const float* __restrict__ radiance,
                         const int32_t* __restrict__ cam, float* __restrict__ rgb_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* hr = H + ray * D * 3;
  const float* rad = radiance + (int64_t)(cam ? cam[ray] : 0) * D * 3;
  // lane l owns flat elements l, l + 96, l + 192, ... of the three interleaved phases: element e = 3 j + c has channel e % 3;
  // with a stride of 96 = 32 * 3 the channel of a lane's elements never changes within a phase
  float acc[3] = {0.f, 0.f, 0.f};
  const int n = D * 3;
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) {
    float a = 0.f;
    for (int e = ph * 32 + lane; e < n; e += 96) a = fmaf(__ldcs(hr + e), __ldg(rad + e), a);
    const int ch = (ph * 32 + lane) % 3;
    acc[0] += ch == 0 ? a : 0.f; acc[1] += ch == 1 ? a : 0.f; acc[2] += ch == 2 ? a : 0.f;
  }
  const float o0 = warp_sum(acc[0]), o1 = warp_sum(acc[1]), o2 = warp_sum(acc[2]);
  if (lane == 0) { rgb_lin[ray * 3] = o0; rgb_lin[ray * 3 + 1] = o1; rgb_lin[ray * 3 + 2] = o2; }
}

// Several illuminations per pass over the collapsed cache: H is read once for NL radiance tables (an illumination sweep is bound by
// streaming H, 12 D bytes per ray and pass).  radiance [NL, D, 3], rgb_lin [NL, R, 3].
template <int NL>
__global__ void __launch_bounds__(LP_WARPS * 32)
relight_collapsed_multi_kernel(const float* __restrict__ H, int64_t R, int D, const float* __restrict__ radiance, float* __restrict__ rgb_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* hr = H + ray * D * 3;
  const int n = D * 3;
  float acc[NL][3];
#pragma unroll
  for (int l = 0; l < NL; ++l) acc[l][0] = acc[l][1] = acc[l][2] = 0.f;
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) {
    float a[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) a[l] = 0.f;
    for (int e = ph * 32 + lane; e < n; e += 96) {
      const float h = __ldcs(hr + e);
#pragma unroll
      for (int l = 0; l < NL; ++l) a[l] = fmaf(h, __ldg(radiance + (size_t)l * n + e), a[l]);
    }
    const int ch = (ph * 32 + lane) % 3;
#pragma unroll
    for (int l = 0; l < NL; ++l) { acc[l][0] += ch == 0 ? a[l] : 0.f; acc[l][1] += ch == 1 ? a[l] : 0.f; acc[l][2] += ch == 2 ? a[l] : 0.f; }
  }
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const float o0 = warp_sum(acc[l][0]), o1 = warp_sum(acc[l][1]), o2 = warp_sum(acc[l][2]);
    if (lane == 0) {
      float* o = rgb_lin + ((size_t)l * R + ray) * 3;
      o[0] = o0; o[1] = o1; o[2] = o2;
    }
  }
}

// Backward of lambert_relight_kernel for a cotangent g [R,3] of rgb_lin:
//   d wa [R,S,3], d normals [R,S,3] (lanes = samples), d vis_sel [R,Dp], d radiance [K,D,3] (lanes = directions, atomics).
// The count of positively lit directions is piecewise constant (no gradient), as in torch (renderers.py:101-106).
__global__ void __launch_bounds__(LP_WARPS * 32)
lambert_relight_bwd_kernel(const float* __restrict__ normals, const float* __restrict__ wa, const float* __restrict__ inv_count,
                           int64_t R, int S, const float* __restrict__ dirs, const int32_t* __restrict__ sel_index, int D, int Dp,
                           const float* __restrict__ radiance, const int32_t* __restrict__ cam, const float* __restrict__ vis_sel,
                           float unocc_vis, const float* __restrict__ g_rgb, float* __restrict__ d_wa, float* __restrict__ d_normals,
                           float* __restrict__ d_vis_sel, float* __restrict__ d_radiance) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * LP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const int64_t krow = (int64_t)(cam ? cam[ray] : 0) * D * 3;
  const float* rad = radiance + krow;
  const float* vr = vis_sel + ray * Dp;
  const float g0 = g_rgb[ray * 3], g1 = g_rgb[ray * 3 + 1], g2 = g_rgb[ray * 3 + 2];
  // ---- pass A: lanes over directions -> d vis, d radiance -------------------------------------------------------
  for (int j = lane; j < D; j += 32) {
    const float lx = dirs[j * 3], ly = dirs[j * 3 + 1], lz = dirs[j * 3 + 2];
    const int sj = sel_index[j];
    const float v = sj >= 0 ? vr[sj] : unocc_vis;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int s = 0; s < S; ++s) {
      const int64_t i = ray * S + s;
      float c = normals[i * 3] * lx + normals[i * 3 + 1] * ly + normals[i * 3 + 2] * lz;
      c = fminf(fmaxf(c, 0.f), 1.f) * inv_count[i];
      c0 = fmaf(wa[i * 3], c, c0); c1 = fmaf(wa[i * 3 + 1], c, c1); c2 = fmaf(wa[i * 3 + 2], c, c2);
    }
    if (sj >= 0 && d_vis_sel) d_vis_sel[ray * Dp + sj] = g0 * c0 * rad[j * 3] + g1 * c1 * rad[j * 3 + 1] + g2 * c2 * rad[j * 3 + 2];
    if (d_radiance) {
      atomicAdd(d_radiance + krow + j * 3, g0 * c0 * v);
      atomicAdd(d_radiance + krow + j * 3 + 1, g1 * c1 * v);
      atomicAdd(d_radiance + krow + j * 3 + 2, g2 * c2 * v);
    }
  }
  // ---- pass B: lanes over samples -> d wa, d normals ------------------------------------------------------------
  for (int s = lane; s < S; s += 32) {
    const int64_t i = ray * S + s;
    const float nx = normals[i * 3], ny = normals[i * 3 + 1], nz = normals[i * 3 + 2];
    const float w0 = wa[i * 3], w1 = wa[i * 3 + 1], w2 = wa[i * 3 + 2];
    const float ic = inv_count[i];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, dn0 = 0.f, dn1 = 0.f, dn2 = 0.f;
    for (int j = 0; j < D; ++j) {
      const float lx = dirs[j * 3], ly = dirs[j * 3 + 1], lz = dirs[j * 3 + 2];
      const int sj = sel_index[j];
      const float v = (sj >= 0 ? vr[sj] : unocc_vis) * ic;
      const float craw = nx * lx + ny * ly + nz * lz;
      const float c = fminf(fmaxf(craw, 0.f), 1.f);
      const float r0 = rad[j * 3] * g0, r1 = rad[j * 3 + 1] * g1, r2 = rad[j * 3 + 2] * g2;
      a0 = fmaf(r0, c * v, a0); a1 = fmaf(r1, c * v, a1); a2 = fmaf(r2, c * v, a2);
      if (craw >= 0.f && craw <= 1.f) {                     // clamp passes the gradient inside [0,1]
        const float k = (w0 * r0 + w1 * r1 + w2 * r2) * v;
        dn0 = fmaf(k, lx, dn0); dn1 = fmaf(k, ly, dn1); dn2 = fmaf(k, lz, dn2);
      }
    }
    d_wa[i * 3] = a0; d_wa[i * 3 + 1] = a1; d_wa[i * 3 + 2] = a2;
    d_normals[i * 3] = dn0; d_normals[i * 3 + 1] = dn1; d_normals[i * 3 + 2] = dn2;
  }
}

// d rgb_lin, d bg, d acc of rgb = srgb(rgb_lin + bg (1 - acc)) (training: linear_to_sRGB clamps to [0,1], no gradient outside)
__global__ void shade_finalize_bwd_kernel(const float* __restrict__ rgb_lin, const float* __restrict__ bg, const float* __restrict__ acc,
                                          const float* __restrict__ g_rgb, int64_t R, float* __restrict__ d_lin, float* __restrict__ d_bg,
                                          float* __restrict__ d_acc) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float dacc = 0.f;
  const float t = 1.0f - acc[r];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = rgb_lin[r * 3 + c] + bg[r * 3 + c] * t;
    float d;
    if (x <= 0.0031308f) d = 12.92f;
    else d = 1.055f / 2.4f * powf(fabsf(x), 1.0f / 2.4f - 1.0f);
    const float y = (x <= 0.0031308f) ? 12.92f * x : 1.055f * powf(fabsf(x), 1.0f / 2.4f) - 0.055f;
    if (!(y >= 0.f && y <= 1.f)) d = 0.f;
    const float gx = g_rgb[r * 3 + c] * d;
    d_lin[r * 3 + c] = gx;
    d_bg[r * 3 + c] = gx * t;
    dacc -= gx * bg[r * 3 + c];
  }
  d_acc[r] = dacc;
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
  if (S >= 32 && cam == nullptr && (size_t)D * 32 <= 96 * 1024) {
    // one thread per sample (full renders); rgb_lin is accumulated with red.global.add, so it is cleared first
    const int64_t N = R * S;
    const int64_t nb = (N + nsk::LPS_THREADS - 1) / nsk::LPS_THREADS;
    NSK_REQUIRE(nb < (1ll << 31), "nsk_lambert_prep: too many samples for one launch");
    const size_t smem = (size_t)D * 2 * sizeof(float4);
    if (smem > 48 * 1024) {
      static nsk::DeviceOnce once;
      if (int err = nsk::device_once(once, "nsk_lambert_prep: shared memory opt-in", nullptr, [] {
            return cudaFuncSetAttribute(nsk::lambert_prep_samples_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
          }))
        return err;
    }
    if (cudaMemsetAsync(rgb_lin, 0, (size_t)R * 3 * sizeof(float), nsk::as_stream(stream)) != cudaSuccess) return nsk::fail("nsk_lambert_prep", "memset");
    nsk::lambert_prep_samples_kernel<<<(unsigned)nb, nsk::LPS_THREADS, smem, nsk::as_stream(stream)>>>(
        normals, wa, N, S, dirs, ddf_mask, D, radiance, unoccluded_vis, inv_count, rgb_lin);
    return nsk::check_launch("lambert_prep_samples_kernel");
  }
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

extern "C" int nsk_lambert_collapse(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                                    const float* dirs, const int32_t* sel_index, int D, int Dp, const float* vis_sel,
                                    float unoccluded_vis, float* H, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1 && D >= 1, "nsk_lambert_collapse: S and D must be >= 1");
  NSK_REQUIRE(normals && wa && inv_count && dirs && sel_index && H && (vis_sel || Dp == 0), "nsk_lambert_collapse: null pointer");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_lambert_collapse: too many rays for one launch");
  nsk::lambert_collapse_kernel<<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      normals, wa, inv_count, R, S, dirs, sel_index, D, Dp, vis_sel, unoccluded_vis, H);
  return nsk::check_launch("lambert_collapse_kernel");
}

extern "C" int nsk_lambert_collapse_sel(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                                        const float* dirs_sel, int Dp, float* G, void* stream) {
  if (R == 0 || Dp == 0) return 0;
  NSK_REQUIRE(S >= 1 && S <= 2048, "nsk_lambert_collapse_sel: S out of range");
  NSK_REQUIRE(normals && wa && inv_count && dirs_sel && G, "nsk_lambert_collapse_sel: null pointer");
  NSK_REQUIRE(R < (1ll << 31), "nsk_lambert_collapse_sel: too many rays for one launch");
  const size_t smem = (size_t)S * 2 * sizeof(float4);
  if (smem > 48 * 1024) {
    static nsk::DeviceOnce once;
    if (int err = nsk::device_once(once, "nsk_lambert_collapse_sel: shared memory opt-in", nullptr, [] {
          return cudaFuncSetAttribute(nsk::lambert_collapse_sel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        }))
      return err;
  }
  nsk::lambert_collapse_sel_kernel<<<(unsigned)R, nsk::LCS_THREADS, smem, nsk::as_stream(stream)>>>(normals, wa, inv_count, R, S, dirs_sel, Dp, G);
  return nsk::check_launch("lambert_collapse_sel_kernel");
}

extern "C" int nsk_relight_collapsed(const float* H, int64_t R, int D, const float* radiance, const int32_t* cam, float* rgb_lin,
                                     void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(D >= 1 && H && radiance && rgb_lin, "nsk_relight_collapsed: null pointer / D");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_relight_collapsed: too many rays for one launch");
  nsk::relight_collapsed_kernel<<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, nsk::as_stream(stream)>>>(H, R, D, radiance, cam, rgb_lin);
  return nsk::check_launch("relight_collapsed_kernel");
}

extern "C" int nsk_relight_collapsed_multi(const float* H, int64_t R, int D, const float* radiance, int NL, float* rgb_lin, void* stream) {
  if (R == 0 || NL == 0) return 0;
  NSK_REQUIRE(D >= 1 && NL >= 1 && H && radiance && rgb_lin, "nsk_relight_collapsed_multi: null pointer / sizes");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_relight_collapsed_multi: too many rays for one launch");
  cudaStream_t st = nsk::as_stream(stream);
  const size_t n = (size_t)D * 3;
  int l = 0;
  while (l < NL) {                       // groups of 4, then 2, then 1 illuminations per pass over H
    const float* rad = radiance + (size_t)l * n;
    float* out = rgb_lin + (size_t)l * R * 3;
    if (NL - l >= 4) { nsk::relight_collapsed_multi_kernel<4><<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, st>>>(H, R, D, rad, out); l += 4; }
    else if (NL - l >= 2) { nsk::relight_collapsed_multi_kernel<2><<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, st>>>(H, R, D, rad, out); l += 2; }
    else { nsk::relight_collapsed_multi_kernel<1><<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, st>>>(H, R, D, rad, out); l += 1; }
  }
  return nsk::check_launch("relight_collapsed_multi_kernel");
}

extern "C" int nsk_lambert_relight_bwd(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                                       const float* dirs, const int32_t* sel_index, int D, int Dp, const float* radiance,
                                       const int32_t* cam, const float* vis_sel, float unoccluded_vis, const float* g_rgb_lin,
                                       float* d_wa, float* d_normals, float* d_vis_sel, float* d_radiance, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1 && D >= 1, "nsk_lambert_relight_bwd: S and D must be >= 1");
  NSK_REQUIRE(normals && wa && inv_count && dirs && sel_index && radiance && g_rgb_lin && d_wa && d_normals && (vis_sel || Dp == 0),
              "nsk_lambert_relight_bwd: null pointer");
  const int64_t blocks = (R + nsk::LP_WARPS - 1) / nsk::LP_WARPS;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_lambert_relight_bwd: too many rays for one launch");
  nsk::lambert_relight_bwd_kernel<<<(unsigned)blocks, nsk::LP_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      normals, wa, inv_count, R, S, dirs, sel_index, D, Dp, radiance, cam, vis_sel, unoccluded_vis, g_rgb_lin, d_wa, d_normals,
      d_vis_sel, d_radiance);
  return nsk::check_launch("lambert_relight_bwd_kernel");
}

extern "C" int nsk_shade_finalize_bwd(const float* rgb_lin, const float* bg, const float* acc, const float* g_rgb, int64_t R,
                                      float* d_rgb_lin, float* d_bg, float* d_acc, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(rgb_lin && bg && acc && g_rgb && d_rgb_lin && d_bg && d_acc, "nsk_shade_finalize_bwd: null pointer");
  nsk::shade_finalize_bwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(rgb_lin, bg, acc, g_rgb, R, d_rgb_lin, d_bg, d_acc);
  return nsk::check_launch("shade_finalize_bwd_kernel");
}

// Light-sum shaders: the per-(row, light) sums of
//   mode 0  reni LambertianShader                      ns_reni/reni/model_components/shaders.py:25-70
//   mode 1  reni BlinnPhongShader                      ns_reni/reni/model_components/shaders.py:73-161
//   mode 2  NeuSky RGBBlinnPhongRendererWithVisibility neusky/model_components/renderers.py:179-253 (per-sample radiance)
// forward and backward.  The reference expands normals to [N,M,3] and materialises [N,M,3] products; here one thread owns a
// row (pixel or ray sample), the M light directions sit in shared memory, the radiance table [K,M,3] is indexed by the row's
// camera (warp-uniform in practice -> one broadcast load per light) and nothing of size N x M is written.  Memory-bound on
// the per-row operands; the light loop is ~25 FLOP + one powf per pair.
#include <algorithm>

#include "nsk_common.cuh"

namespace nsk {
namespace shaders {

constexpr int THREADS = 128;

struct Args {
  const float* albedo;     // [N,3]
  const float* normals;    // [N,3]
  const float* specular;   // [N,3]  mode 1
  const float* shininess;  // [N]    modes 1, 2
  const float* view;       // [N,3]  modes 1, 2 (world-space view direction added to every light direction)
  const float* dirs;       // [M,3] shared, or [N,M,3] when dirs_per_row
  const float* radiance;   // [K,M,3]
  const int* cam;          // [N] row -> K index, or NULL (0)
  const float* vis;        // [N / rows_per_vis, M] or NULL (1)   mode 2
  int64_t N;
  int M;
  int rows_per_vis;
  int dirs_per_row;
  int normalize_dirs;      // mode 1 option (shaders.py:117-118)
};

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// Blinn-Phong normalisation (n + 2) / (4 (2 - exp(-n / 2))) (shaders.py:146-148) and its derivative
__device__ __forceinline__ float bp_norm(float s, float* dF) {
  const float E = expf(-0.5f * s);
  const float den = 2.f - E;
  if (dF) *dF = (den - 0.5f * (s + 2.f) * E) / (4.f * den * den);
  return (s + 2.f) / (4.f * den);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS)
shade_fwd_kernel(Args a, float* __restrict__ out_a, float* __restrict__ out_b, const float* __restrict__ weights,
                 float* __restrict__ rgb_lin, int S) {
  extern __shared__ float s_dirs[];
  if (!a.dirs_per_row) {
    for (int t = threadIdx.x; t < a.M * 3; t += blockDim.x) s_dirs[t] = a.dirs[t];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const float n[3] = {a.normals[i * 3], a.normals[i * 3 + 1], a.normals[i * 3 + 2]};
  const float al[3] = {a.albedo[i * 3], a.albedo[i * 3 + 1], a.albedo[i * 3 + 2]};
  float v[3] = {0.f, 0.f, 0.f}, shin = 0.f;
  if (MODE >= 1) {
    v[0] = a.view[i * 3]; v[1] = a.view[i * 3 + 1]; v[2] = a.view[i * 3 + 2];
    shin = a.shininess[i];
  }
  const float* L = a.radiance + (size_t)(a.cam ? a.cam[i] : 0) * a.M * 3;
  const float* vr = (MODE == 2 && a.vis) ? a.vis + (size_t)(i / a.rows_per_vis) * a.M : nullptr;
  const float* dr = a.dirs_per_row ? a.dirs + (size_t)i * a.M * 3 : s_dirs;
  float Sd[3] = {0.f, 0.f, 0.f}, Ss[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < a.M; ++j) {
    float l[3] = {dr[j * 3], dr[j * 3 + 1], dr[j * 3 + 2]};
    if (MODE == 1 && a.normalize_dirs) {
      const float ln = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      l[0] /= ln; l[1] /= ln; l[2] /= ln;
    }
    const float Lj[3] = {__ldg(L + j * 3), __ldg(L + j * 3 + 1), __ldg(L + j * 3 + 2)};
    const float nl = n[0] * l[0] + n[1] * l[1] + n[2] * l[2];
    float c = MODE == 2 ? clamp01(nl) : fmaxf(nl, 0.f);
    float p = 0.f;
    if (MODE >= 1) {
      float h[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
      const float hn = sqrtf(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
      const float inv = 1.f / (MODE == 1 ? hn + 1e-8f : hn);
      const float nh = (n[0] * h[0] + n[1] * h[1] + n[2] * h[2]) * inv;
      p = powf(MODE == 2 ? clamp01(nh) : fmaxf(nh, 0.f), shin);
    }
    if (MODE == 2) {
      const float vj = vr ? vr[j] : 1.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) Sd[ch] += Lj[ch] * vj * (al[ch] * c + p);
    } else {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        Sd[ch] += c * Lj[ch];
        if (MODE == 1) Ss[ch] += p * Lj[ch];
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      out_a[i * 3 + ch] = Sd[ch];
      out_b[i * 3 + ch] = al[ch] * Sd[ch];
    }
  } else if (MODE == 1) {
    const float F = bp_norm(shin, nullptr);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out_a[i * 3 + ch] = fmaxf(al[ch] * Sd[ch] + a.specular[i * 3 + ch] * F * Ss[ch], 1e-3f);
  } else {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out_a[i * 3 + ch] = Sd[ch];
    if (weights && rgb_lin) {     // comp_rgb = sum_s w * radiance (renderers.py:247), reduced with red.global.add like K4
      const float w = weights[i];
      const int64_t r = i / S;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) atomicAdd(rgb_lin + r * 3 + ch, w * Sd[ch]);
    }
  }
}

// Sum over the warp when every lane targets the same address (same camera / same ray), else per-lane atomics.
__device__ __forceinline__ void warp_add(float* addr_base, bool uniform, float val) {
  if (uniform) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((threadIdx.x & 31) == 0 && val != 0.f) atomicAdd(addr_base, val);
  } else if (val != 0.f) {
    atomicAdd(addr_base, val);
  }
}

// g_a: cotangent on out_a [N,3]; g_b: cotangent on out_b (mode 0, may be NULL).  Row gradients are overwritten, d_radiance
// [K,M,3] and d_vis are accumulated (caller zero-fills).
template <int MODE>
__global__ void __launch_bounds__(THREADS)
shade_bwd_kernel(Args a, const float* __restrict__ g_a, const float* __restrict__ g_b, float* __restrict__ d_albedo,
                 float* __restrict__ d_normals, float* __restrict__ d_specular, float* __restrict__ d_shin,
                 float* __restrict__ d_radiance, float* __restrict__ d_vis) {
  extern __shared__ float s_dirs[];
  if (!a.dirs_per_row) {
    for (int t = threadIdx.x; t < a.M * 3; t += blockDim.x) s_dirs[t] = a.dirs[t];
    __syncthreads();
  }
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < a.N;
  const int64_t i = live ? i0 : a.N - 1;       // dead lanes shadow the last row with zero cotangents (keeps warps whole)
  const float n[3] = {a.normals[i * 3], a.normals[i * 3 + 1], a.normals[i * 3 + 2]};
  const float al[3] = {a.albedo[i * 3], a.albedo[i * 3 + 1], a.albedo[i * 3 + 2]};
  float v[3] = {0.f, 0.f, 0.f}, sp[3] = {0.f, 0.f, 0.f}, shin = 0.f, F = 0.f, dF = 0.f;
  if (MODE >= 1) {
    v[0] = a.view[i * 3]; v[1] = a.view[i * 3 + 1]; v[2] = a.view[i * 3 + 2];
    shin = a.shininess[i];
  }
  if (MODE == 1) {
    sp[0] = a.specular[i * 3]; sp[1] = a.specular[i * 3 + 1]; sp[2] = a.specular[i * 3 + 2];
    F = bp_norm(shin, &dF);
  }
  const int k = a.cam ? a.cam[i] : 0;
  const int64_t vrow = i / a.rows_per_vis;
  const float* L = a.radiance + (size_t)k * a.M * 3;
  const float* vr = (MODE == 2 && a.vis) ? a.vis + (size_t)vrow * a.M : nullptr;
  const float* dr = a.dirs_per_row ? a.dirs + (size_t)i * a.M * 3 : s_dirs;
  const bool cam_uniform = __all_sync(0xffffffffu, k == __shfl_sync(0xffffffffu, k, 0));
  const bool vis_uniform = __all_sync(0xffffffffu, vrow == __shfl_sync(0xffffffffu, vrow, 0));

  float g[3] = {0.f, 0.f, 0.f}, gb[3] = {0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      g[ch] = g_a ? g_a[i * 3 + ch] : 0.f;
      gb[ch] = (MODE == 0 && g_b) ? g_b[i * 3 + ch] : 0.f;
    }
  }
  // MODE 1: the output clamp(min=1e-3) passes the cotangent where the unclamped colour is >= 1e-3: needs the forward sums first
  float Sd[3] = {0.f, 0.f, 0.f}, Ss[3] = {0.f, 0.f, 0.f};
  if (MODE == 1) {
    for (int j = 0; j < a.M; ++j) {
      float l[3] = {dr[j * 3], dr[j * 3 + 1], dr[j * 3 + 2]};
      if (a.normalize_dirs) {
        const float ln = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
        l[0] /= ln; l[1] /= ln; l[2] /= ln;
      }
      const float c = fmaxf(n[0] * l[0] + n[1] * l[1] + n[2] * l[2], 0.f);
      float h[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
      const float inv = 1.f / (sqrtf(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]) + 1e-8f);
      const float p = powf(fmaxf((n[0] * h[0] + n[1] * h[1] + n[2] * h[2]) * inv, 0.f), shin);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float Lc = __ldg(L + j * 3 + ch);
        Sd[ch] += c * Lc;
        Ss[ch] += p * Lc;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (!(al[ch] * Sd[ch] + sp[ch] * F * Ss[ch] >= 1e-3f)) g[ch] = 0.f;
  }
  // effective cotangents on the diffuse sum (G) and on the specular sum (Gs), per channel
  float G[3], Gs[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    G[ch] = MODE == 0 ? g[ch] + gb[ch] * al[ch] : g[ch] * al[ch];
    Gs[ch] = MODE == 1 ? g[ch] * sp[ch] * F : g[ch];
  }
  float dA[3] = {0.f, 0.f, 0.f}, dN[3] = {0.f, 0.f, 0.f}, dS = 0.f, sd0[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < a.M; ++j) {
    float l[3] = {dr[j * 3], dr[j * 3 + 1], dr[j * 3 + 2]};
    if (MODE == 1 && a.normalize_dirs) {
      const float ln = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      l[0] /= ln; l[1] /= ln; l[2] /= ln;
    }
    const float Lj[3] = {__ldg(L + j * 3), __ldg(L + j * 3 + 1), __ldg(L + j * 3 + 2)};
    const float nl = n[0] * l[0] + n[1] * l[1] + n[2] * l[2];
    const float c = MODE == 2 ? clamp01(nl) : fmaxf(nl, 0.f);
    const bool c_open = MODE == 2 ? (nl >= 0.f && nl <= 1.f) : nl >= 0.f;       // torch clamp backward: inclusive bounds
    float p = 0.f, nh = 0.f, hu[3] = {0.f, 0.f, 0.f};
    bool p_open = false;
    if (MODE >= 1) {
      float h[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
      const float inv = 1.f / (sqrtf(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]) + (MODE == 1 ? 1e-8f : 0.f));
      hu[0] = h[0] * inv; hu[1] = h[1] * inv; hu[2] = h[2] * inv;
      const float raw = n[0] * hu[0] + n[1] * hu[1] + n[2] * hu[2];
      nh = MODE == 2 ? clamp01(raw) : fmaxf(raw, 0.f);
      p_open = MODE == 2 ? (raw >= 0.f && raw <= 1.f) : raw >= 0.f;
      p = powf(nh, shin);
    }
    const float vj = (MODE == 2 && vr) ? vr[j] : 1.f;
    const float gL = G[0] * Lj[0] + G[1] * Lj[1] + G[2] * Lj[2];              // sum_c G_c L_jc
    const float gsL = MODE >= 1 ? Gs[0] * Lj[0] + Gs[1] * Lj[1] + Gs[2] * Lj[2] : 0.f;
    if (MODE == 2) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) dA[ch] += g[ch] * Lj[ch] * vj * c;
    } else {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) sd0[ch] += c * Lj[ch];
    }
    if (c_open) {
      const float f = gL * vj;
      dN[0] += f * l[0]; dN[1] += f * l[1]; dN[2] += f * l[2];
    }
    if (MODE >= 1) {
      if (p_open && nh > 0.f) {
        const float f = gsL * vj * shin * powf(nh, shin - 1.f);
        dN[0] += f * hu[0]; dN[1] += f * hu[1]; dN[2] += f * hu[2];
        dS += gsL * vj * p * logf(nh);
      }
    }
    if (d_radiance) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float val;
        if (MODE == 0) val = G[ch] * c;
        else if (MODE == 1) val = g[ch] * (al[ch] * c + sp[ch] * F * p);
        else val = g[ch] * vj * (al[ch] * c + p);
        warp_add(d_radiance + ((size_t)k * a.M + j) * 3 + ch, cam_uniform, val);
      }
    }
    if (MODE == 2 && d_vis) {
      float val = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) val += g[ch] * Lj[ch] * (al[ch] * c + p);
      warp_add(d_vis + (size_t)vrow * a.M + j, vis_uniform, val);
    }
  }
  if (!live) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (d_albedo) d_albedo[i * 3 + ch] = MODE == 0 ? gb[ch] * sd0[ch] : (MODE == 1 ? g[ch] * sd0[ch] : dA[ch]);
    if (d_normals) d_normals[i * 3 + ch] = dN[ch];
    if (MODE == 1 && d_specular) d_specular[i * 3 + ch] = g[ch] * F * Ss[ch];
  }
  if (MODE >= 1 && d_shin) {
    float extra = 0.f;
    if (MODE == 1) extra = dF * (g[0] * sp[0] * Ss[0] + g[1] * sp[1] * Ss[1] + g[2] * sp[2] * Ss[2]);
    d_shin[i] = dS + extra;
  }
}

static int check_args(const Args& a, int mode, const char* who) {
  if (!(a.albedo && a.normals && a.dirs && a.radiance)) return fail(who, "null pointer");
  if (mode >= 1 && !(a.shininess && a.view)) return fail(who, "shininess / view_dirs required for Blinn-Phong");
  if (mode == 1 && !a.specular) return fail(who, "specular required");
  if (a.M <= 0 || a.rows_per_vis <= 0) return fail(who, "M / rows_per_vis");
  if (!a.dirs_per_row && (size_t)a.M * 3 * sizeof(float) > 200 * 1024) return fail(who, "too many light directions for shared memory");
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess;
  return 0;
}

}  // namespace shaders
}  // namespace nsk

using namespace nsk;
using namespace nsk::shaders;

extern "C" int nsk_shade_lights_fwd(int mode, const float* albedo, const float* normals, const float* specular,
                                    const float* shininess, const float* view_dirs, const float* dirs, int dirs_per_row,
                                    int normalize_dirs, const float* radiance, const int* cam, const float* vis,
                                    int rows_per_vis, int64_t N, int M, float* out_a, float* out_b, const float* weights,
                                    float* rgb_lin, int S, void* stream) {
  Args a{albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis, N, M, rows_per_vis, dirs_per_row, normalize_dirs};
  NSK_REQUIRE(mode >= 0 && mode <= 2, "nsk_shade_lights_fwd: mode");
  if (check_args(a, mode, "nsk_shade_lights_fwd")) return 1;
  NSK_REQUIRE(out_a && (mode != 0 || out_b), "nsk_shade_lights_fwd: null output");
  NSK_REQUIRE(!(weights && rgb_lin) || S > 0, "nsk_shade_lights_fwd: S");
  if (N == 0) return 0;
  const size_t smem = dirs_per_row ? 0 : (size_t)M * 3 * sizeof(float);
  const unsigned grid = (unsigned)((N + THREADS - 1) / THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    if (set_smem(shade_fwd_kernel<0>, smem)) return fail("nsk_shade_lights_fwd", "shared memory opt-in");
    shade_fwd_kernel<0><<<grid, THREADS, smem, st>>>(a, out_a, out_b, nullptr, nullptr, 1);
  } else if (mode == 1) {
    if (set_smem(shade_fwd_kernel<1>, smem)) return fail("nsk_shade_lights_fwd", "shared memory opt-in");
    shade_fwd_kernel<1><<<grid, THREADS, smem, st>>>(a, out_a, out_b, nullptr, nullptr, 1);
  } else {
    if (set_smem(shade_fwd_kernel<2>, smem)) return fail("nsk_shade_lights_fwd", "shared memory opt-in");
    shade_fwd_kernel<2><<<grid, THREADS, smem, st>>>(a, out_a, out_b, weights, rgb_lin, S > 0 ? S : 1);
  }
  return check_launch("shade_fwd_kernel");
}

extern "C" int nsk_shade_lights_bwd(int mode, const float* albedo, const float* normals, const float* specular,
                                    const float* shininess, const float* view_dirs, const float* dirs, int dirs_per_row,
                                    int normalize_dirs, const float* radiance, const int* cam, const float* vis,
                                    int rows_per_vis, int64_t N, int M, const float* g_a, const float* g_b, float* d_albedo,
                                    float* d_normals, float* d_specular, float* d_shininess, float* d_radiance,
                                    float* d_vis, void* stream) {
  Args a{albedo, normals, specular, shininess, view_dirs, dirs, radiance, cam, vis, N, M, rows_per_vis, dirs_per_row, normalize_dirs};
  NSK_REQUIRE(mode >= 0 && mode <= 2, "nsk_shade_lights_bwd: mode");
  if (check_args(a, mode, "nsk_shade_lights_bwd")) return 1;
  NSK_REQUIRE(g_a || (mode == 0 && g_b), "nsk_shade_lights_bwd: no cotangent");
  if (N == 0) return 0;
  const size_t smem = dirs_per_row ? 0 : (size_t)M * 3 * sizeof(float);
  const unsigned grid = (unsigned)((N + THREADS - 1) / THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    if (set_smem(shade_bwd_kernel<0>, smem)) return fail("nsk_shade_lights_bwd", "shared memory opt-in");
    shade_bwd_kernel<0><<<grid, THREADS, smem, st>>>(a, g_a, g_b, d_albedo, d_normals, d_specular, d_shininess, d_radiance, d_vis);
  } else if (mode == 1) {
    if (set_smem(shade_bwd_kernel<1>, smem)) return fail("nsk_shade_lights_bwd", "shared memory opt-in");
    shade_bwd_kernel<1><<<grid, THREADS, smem, st>>>(a, g_a, g_b, d_albedo, d_normals, d_specular, d_shininess, d_radiance, d_vis);
  } else {
    if (set_smem(shade_bwd_kernel<2>, smem)) return fail("nsk_shade_lights_bwd", "shared memory opt-in");
    shade_bwd_kernel<2><<<grid, THREADS, smem, st>>>(a, g_a, g_b, d_albedo, d_normals, d_specular, d_shininess, d_radiance, d_vis);
  }
  return check_launch("shade_bwd_kernel");
}

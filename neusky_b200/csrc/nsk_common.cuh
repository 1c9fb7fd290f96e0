// Shared device/host helpers for the neusky_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>

#include "../../include/neusky_b200.h"

namespace nsk {

// ---- thread-local error string behind nsk_last_error() ------------------------------------
char* last_error_buffer();  // defined in abi.cu
inline int fail(const char* what, const char* detail = "") {
  snprintf(last_error_buffer(), 512, "%s%s%s", what, detail[0] ? ": " : "", detail);
  return 1;
}
inline int check_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(kernel, cudaGetErrorString(e));
  return 0;
}
#define NSK_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return nsk::fail(msg, #cond); \
  } while (0)

// ---- DDF network shape (neusky/configs/neusky_config.py:162-177) ---------------------------
constexpr int DDF_HID = 256;      // hidden_features == mapping_features
constexpr int DDF_LAYERS = 5;     // hidden_layers == mapping_layers
constexpr int DDF_LEVELS = 16;    // hash levels (directional_distance_field.py:140)
constexpr int DDF_MAP_IN = 35;    // 3 + 16*2
constexpr int DDF_DIR_IN = 15;    // 3 + 12
constexpr int DDF_FILM = DDF_LAYERS * DDF_HID * 2;  // 2560

// ---- hash grid (nerfstudio torch semantics) ------------------------------------
__device__ __forceinline__ uint32_t hash3(int cx, int cy, int cz, uint32_t mask) {
  // low log2(T) bits of the int64 product-XOR == uint32 wrap-around product-XOR
  return ((uint32_t)cx ^ ((uint32_t)cy * 2654435761u) ^ ((uint32_t)cz * 805459861u)) & mask;
}

// Corner order 0..7 = (c,c,c),(c,f,c),(f,f,c),(f,c,c),(c,c,f),(c,f,f),(f,f,f),(f,c,f)
__device__ __forceinline__ void hash_corners(float sx, float sy, float sz, uint32_t mask, uint32_t idx[8],
                                             float& ox, float& oy, float& oz) {
  const float fxf = floorf(sx), fyf = floorf(sy), fzf = floorf(sz);
  const int fx = (int)fxf, fy = (int)fyf, fz = (int)fzf;
  const int cx = (int)ceilf(sx), cy = (int)ceilf(sy), cz = (int)ceilf(sz);
  ox = __fsub_rn(sx, fxf);
  oy = __fsub_rn(sy, fyf);
  oz = __fsub_rn(sz, fzf);
  idx[0] = hash3(cx, cy, cz, mask);
  idx[1] = hash3(cx, fy, cz, mask);
  idx[2] = hash3(fx, fy, cz, mask);
  idx[3] = hash3(fx, cy, cz, mask);
  idx[4] = hash3(cx, cy, fz, mask);
  idx[5] = hash3(cx, fy, fz, mask);
  idx[6] = hash3(fx, fy, fz, mask);
  idx[7] = hash3(fx, cy, fz, mask);
}

// ---- grid semantics switch (SURVEY A.3 "tcnn differences", neusky_b200/tcnn_import.py) ------------------------------------------------
// meta == nullptr: the nerfstudio torch HashEncoding the rest of this file implements (scale from `scalings`, corners floor / ceil,
// linear weights, prime hash & mask).  meta != nullptr: tiny-cuda-nn's grid as imported from a reference checkpoint: per level
// (float bits of scale, resolution, size, dense) -- pos = x * scale + 0.5, corners floor / floor + 1, dense index x + y res + z res^2
// or the same prime hash, both modulo the level's size, linear or smoothstep weights.  The table keeps our [L * T, F] layout.
struct GridMode {
  const int4* meta;
  int smoothstep;
};
// Corners + interpolation weights of level `lev` at the (already normalised) position p.  `ox, oy, oz` are the weights of the UPPER corner
// per axis in the corner order of hash_corners() (so hash_interp / hash_interp_grad apply unchanged); `dw` = d weight / d (grid
// position) per axis (1 for linear weights) and `scale` = d (grid position) / d p: the chain rule factors of the analytic normal.
__device__ __forceinline__ void grid_corners(const GridMode gm, int lev, float px, float py, float pz, float ns_scale, uint32_t mask, uint32_t idx[8],
                                             float& ox, float& oy, float& oz, float (&dw)[3], float& scale) {
  if (gm.meta == nullptr) {
    hash_corners(__fmul_rn(px, ns_scale), __fmul_rn(py, ns_scale), __fmul_rn(pz, ns_scale), mask, idx, ox, oy, oz);
    dw[0] = dw[1] = dw[2] = 1.0f;
    scale = ns_scale;
    return;
  }
  const int4 m = __ldg(gm.meta + lev);
  scale = __int_as_float(m.x);
  const uint32_t res = (uint32_t)m.y, size = (uint32_t)m.z;
  const bool dense = m.w != 0;
  const float p[3] = {px, py, pz};
  uint32_t g[3];
  float w[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(p[d], scale, 0.5f);
    const float fl = floorf(pos);
    g[d] = (uint32_t)(int)fl;
    const float t = pos - fl;
    w[d] = gm.smoothstep ? t * t * (3.0f - 2.0f * t) : t;
    dw[d] = gm.smoothstep ? 6.0f * t * (1.0f - t) : 1.0f;
  }
  ox = w[0]; oy = w[1]; oz = w[2];
  // corner order of hash_corners(): "c" = upper corner (g + 1), "f" = lower corner (g)
  const int ux[8] = {1, 1, 0, 0, 1, 1, 0, 0}, uy[8] = {1, 0, 0, 1, 1, 0, 0, 1}, uz[8] = {1, 1, 1, 1, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t gx = g[0] + ux[c], gy = g[1] + uy[c], gz = g[2] + uz[c];
    idx[c] = (dense ? gx + gy * res + gz * res * res : (gx ^ (gy * 2654435761u) ^ (gz * 805459861u))) % size;
  }
}

// Interpolation in the oracle's operation order, without FMA contraction (bit-exact vs torch).
__device__ __forceinline__ float lerp_ns(float a, float b, float o) {
  // a*o + b*(1-o)
  return __fadd_rn(__fmul_rn(a, o), __fmul_rn(b, __fsub_rn(1.0f, o)));
}
__device__ __forceinline__ float2 hash_interp(const float2 f[8], float ox, float oy, float oz) {
  float2 r;
  {
    float f03 = lerp_ns(f[0].x, f[3].x, ox), f12 = lerp_ns(f[1].x, f[2].x, ox);
    float f56 = lerp_ns(f[5].x, f[6].x, ox), f47 = lerp_ns(f[4].x, f[7].x, ox);
    float f0312 = lerp_ns(f03, f12, oy), f4756 = lerp_ns(f47, f56, oy);
    r.x = lerp_ns(f0312, f4756, oz);
  }
  {
    float f03 = lerp_ns(f[0].y, f[3].y, ox), f12 = lerp_ns(f[1].y, f[2].y, ox);
    float f56 = lerp_ns(f[5].y, f[6].y, ox), f47 = lerp_ns(f[4].y, f[7].y, ox);
    float f0312 = lerp_ns(f03, f12, oy), f4756 = lerp_ns(f47, f56, oy);
    r.y = lerp_ns(f0312, f4756, oz);
  }
  return r;
}

// ---- geometry of a (surface point, light direction) pair ------------------------------------
// neusky/models/neusky_model.py:1590-1622: exit point of the ray p + t*l on the sphere |x| = r.
__device__ __forceinline__ void sphere_exit(const float p[3], const float l_in[3], float radius, float q[3], float& t) {
  const float ln = sqrtf(l_in[0] * l_in[0] + l_in[1] * l_in[1] + l_in[2] * l_in[2]);
  const float l[3] = {l_in[0] / ln, l_in[1] / ln, l_in[2] / ln};
  const float b = 2.0f * (l[0] * p[0] + l[1] * p[1] + l[2] * p[2]);
  const float c = (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) - radius * radius;
  const float disc = fmaxf(b * b - 4.0f * c, 0.0f);
  const float sq = sqrtf(disc);
  t = fmaxf((-b - sq) * 0.5f, (-b + sq) * 0.5f);
  q[0] = p[0] + t * l[0];
  q[1] = p[1] + t * l[1];
  q[2] = p[2] + t * l[2];
}

// neusky/models/ddf_model.py:158-200: direction expressed in the local frame of sphere point q
// (y = -q, x = normalize(up x y), z = normalize(y x x)); d_local = M^T d with columns (x,y,z).
__device__ __forceinline__ void ddf_local_dir(const float q[3], const float d[3], float out[3]) {
  const float y[3] = {-q[0], -q[1], -q[2]};
  // up = (0,0,1): up x y = (-y1, y0, 0)
  float x[3] = {-y[1], y[0], 0.0f};
  const float xn = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  x[0] /= xn; x[1] /= xn; x[2] /= xn;
  float z[3] = {y[1] * x[2] - y[2] * x[1], y[2] * x[0] - y[0] * x[2], y[0] * x[1] - y[1] * x[0]};
  const float zn = sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
  z[0] /= zn; z[1] /= zn; z[2] /= zn;
  out[0] = x[0] * d[0] + x[1] * d[1] + x[2] * d[2];
  out[1] = y[0] * d[0] + y[1] * d[1] + y[2] * d[2];
  out[2] = z[0] * d[0] + z[1] * d[1] + z[2] * d[2];
}

// DDF direction features (directional_distance_field.py:270-271 with NeRFEncoding(3, 2 freqs {1,4},
// no input) [NS-mem A.2]): [d(3), sin(2pi d f) (6, index dim*2+f), sin(2pi d f + pi/2) (6)] = 15.
__device__ __forceinline__ void ddf_dir_features(const float dl[3], float feat[15]) {
  const float TWO_PI = 6.283185307179586f;
  const float HALF_PI = 1.5707963267948966f;
  feat[0] = dl[0]; feat[1] = dl[1]; feat[2] = dl[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float s = TWO_PI * dl[d];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const float a = s * (f == 0 ? 1.0f : 4.0f);
      feat[3 + d * 2 + f] = sinf(a);
      feat[9 + d * 2 + f] = sinf(a + HALF_PI);
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// visibility of one pair given the DDF output and the pair geometry (neusky_model.py:1724-1740)
__device__ __forceinline__ float visibility_from_ddf(float ddf_dist, float term_dist, float radius, float thr, float scale) {
  const float gt = fminf(term_dist, 2.0f * radius);
  return 1.0f - sigmoidf_(scale * ((gt - ddf_dist) - thr));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- per-device one-time setup ----------------------------------------------------------------
// cudaFuncSetAttribute (dynamic shared-memory opt-in) and the SM count are properties of ONE device; a process may
// drive several devices and several host threads (training thread + viewer thread on another GPU).  Each launch site
// owns one `static DeviceOnce` table, indexed by the CURRENT device; the first call on every device runs `setup`.
// Racing threads may both run `setup` (idempotent); nothing here is a per-process "configured" flag.
constexpr int MAX_DEVICES = 64;
struct DeviceOnce {
  std::atomic<int> ready[MAX_DEVICES];
  int num_sms[MAX_DEVICES];
};
template <typename F>
inline int device_once(DeviceOnce& st, const char* what, int* num_sms, F&& setup) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));
  if (dev < 0 || dev >= MAX_DEVICES) return fail(what, "device index out of range");
  if (!st.ready[dev].load(std::memory_order_acquire)) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = setup();
    if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));
    st.num_sms[dev] = n > 0 ? n : 148;
    st.ready[dev].store(1, std::memory_order_release);
  }
  if (num_sms) *num_sms = st.num_sms[dev];
  return 0;
}

}  // namespace nsk

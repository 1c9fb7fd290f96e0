// K1: multiresolution hash-grid encode, forward / backward / index dump.
// Semantics: nerfstudio HashEncoding.pytorch_fwd [NS-mem, SURVEY A.3] -- the substitution for
// tcnn.Encoding at neusky/fields/sdf_albedo_field.py:117-130 and
// neusky/fields/directional_distance_field.py:139-156.
//
// Layout / mapping: a warp owns 32 consecutive points (lane = point) and walks the levels, so all
// 32 lanes gather from the same level's table slice at once (coarse levels stay L1/L2 resident).
// Each lane issues the 8 corner gathers of a level as independent 8-byte loads (two levels in
// flight = 16 outstanding loads per lane).  Results are staged in shared memory and written back
// as one contiguous 4 KB run per warp (32 points x 128 B) with 16-byte stores.
// HBM-bound roofline: 12 B in + 16*8*8 B gathers + 128 B out = 1164 B per point.
#include "nsk_common.cuh"

namespace nsk {

constexpr int HE_WARPS = 8;
constexpr int HE_MAX_LEVELS = 16;

__global__ void __launch_bounds__(HE_WARPS * 32)
hash_encode_fwd_kernel(const float* __restrict__ x, int64_t n, const float2* __restrict__ table,
                       const float* __restrict__ scalings, int L, int log2_T, float* __restrict__ out) {
  __shared__ float s_scale[HE_MAX_LEVELS];
  __shared__ __align__(16) float stage[HE_WARPS][32][2 * HE_MAX_LEVELS + 1];  // +1: conflict-free column reads
  if (threadIdx.x < L) s_scale[threadIdx.x] = scalings[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t mask = (1u << log2_T) - 1u;
  const int64_t warps_total = (int64_t)gridDim.x * HE_WARPS;
  const int F = 2 * L;
  for (int64_t base = ((int64_t)blockIdx.x * HE_WARPS + warp) * 32; base < n; base += warps_total * 32) {
    const int64_t p = base + lane;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (p < n) { px = x[p * 3 + 0]; py = x[p * 3 + 1]; pz = x[p * 3 + 2]; }
    for (int l = 0; l < L; l += 2) {
      float2 f[2][8];
      float ox[2], oy[2], oz[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int lev = min(l + u, L - 1);
        const float s = s_scale[lev];
        uint32_t idx[8];
        hash_corners(__fmul_rn(px, s), __fmul_rn(py, s), __fmul_rn(pz, s), mask, idx, ox[u], oy[u], oz[u]);
        const float2* tl = table + ((size_t)lev << log2_T);
#pragma unroll
        for (int c = 0; c < 8; ++c) f[u][c] = __ldg(tl + idx[c]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (l + u < L) {
          const float2 r = hash_interp(f[u], ox[u], oy[u], oz[u]);
          stage[warp][lane][2 * (l + u) + 0] = r.x;
          stage[warp][lane][2 * (l + u) + 1] = r.y;
        }
      }
    }
    __syncwarp();
    // coalesced write-back: the warp's 32 x F floats are contiguous in `out`
    const int64_t valid = min((int64_t)32, n - base);
    float* dst = out + base * F;
    for (int i = lane; i < (int)valid * F; i += 32) dst[i] = stage[warp][i / F][i % F];
    __syncwarp();
  }
}

// Backward (scatter): thread = point, a warp = 32 consecutive points walking the levels together, so all in-flight
// reductions of a warp land in one level's table slice (coarse levels: a handful of L2 lines).  Each thread reads its
// own contiguous 8 L-byte gradient row exactly once (16-byte loads, every fetched sector fully used) and x once.
// The first version ran one thread per (point, level) with the level in blockIdx.y: every level re-read a 32-byte
// sector of grad_out for 8 useful bytes, 18 GB of DRAM reads per 7.9 M points (profiles/r02_ncu_hbm_kernels_summary.txt).
__global__ void __launch_bounds__(256)
hash_encode_bwd_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ scalings, int L,
                       int log2_T, const float* __restrict__ grad_out, float* __restrict__ grad_table) {
  __shared__ float s_scale[HE_MAX_LEVELS];
  if (threadIdx.x < L) s_scale[threadIdx.x] = scalings[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t mask = (1u << log2_T) - 1u;
  const float px = x[i * 3], py = x[i * 3 + 1], pz = x[i * 3 + 2];
  const float* grow = grad_out + i * 2 * L;
  // corner c uses (x: c or f, y: c or f, z: c or f) -> weight index 0 for "c" (offset), 1 for "f"
  const int ux[8] = {0, 0, 1, 1, 0, 0, 1, 1}, uy[8] = {0, 1, 1, 0, 0, 1, 1, 0}, uz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  auto scatter = [&](int lev, float gx, float gy) {
    const float s = s_scale[lev];
    uint32_t idx[8];
    float ox, oy, oz;
    hash_corners(__fmul_rn(px, s), __fmul_rn(py, s), __fmul_rn(pz, s), mask, idx, ox, oy, oz);
    const float wx[2] = {ox, 1.f - ox}, wy[2] = {oy, 1.f - oy}, wz[2] = {oz, 1.f - oz};
    float2* tl = reinterpret_cast<float2*>(grad_table) + ((size_t)lev << log2_T);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = wx[ux[c]] * wy[uy[c]] * wz[uz[c]];
      atomicAdd(tl + idx[c], make_float2(w * gx, w * gy));
    }
  };
  if ((L & 1) == 0) {
    // rows are 8 L bytes: 16-byte aligned when L is even
#pragma unroll 2
    for (int lev = 0; lev < L; lev += 2) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(grow) + (lev >> 1));
      scatter(lev, g.x, g.y);
      scatter(lev + 1, g.z, g.w);
    }
  } else {
    for (int lev = 0; lev < L; ++lev) {
      const float2 g = __ldg(reinterpret_cast<const float2*>(grow) + lev);
      scatter(lev, g.x, g.y);
    }
  }
}

__global__ void hash_indices_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ scalings, int L,
                                    int log2_T, int64_t* __restrict__ idx_out, float* __restrict__ off_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  const int64_t p = i / L;
  const int lev = (int)(i % L);
  const uint32_t mask = (1u << log2_T) - 1u;
  const float s = scalings[lev];
  uint32_t idx[8];
  float ox, oy, oz;
  hash_corners(__fmul_rn(x[p * 3], s), __fmul_rn(x[p * 3 + 1], s), __fmul_rn(x[p * 3 + 2], s), mask, idx, ox, oy, oz);
  for (int c = 0; c < 8; ++c) idx_out[i * 8 + c] = (int64_t)idx[c] + ((int64_t)lev << log2_T);
  off_out[i * 3 + 0] = ox; off_out[i * 3 + 1] = oy; off_out[i * 3 + 2] = oz;
}

// ---- first derivative wrt the position and its own backward (double backward of the encode) ---------------------
// coefficient of corner value f[c] in d interp / d o_axis (corner order of hash_corners)
__device__ __forceinline__ void interp_grad_coefs(float ox, float oy, float oz, float cf[3][8]) {
  const float px[2] = {ox, 1.f - ox}, py[2] = {oy, 1.f - oy}, pz[2] = {oz, 1.f - oz};
  const int ux[8] = {0, 0, 1, 1, 0, 0, 1, 1}, uy[8] = {0, 1, 1, 0, 0, 1, 1, 0}, uz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float sx = ux[c] ? -1.f : 1.f, sy = uy[c] ? -1.f : 1.f, sz = uz[c] ? -1.f : 1.f;   // d w / d o: +1 for the "ceil" side
    cf[0][c] = sx * py[uy[c]] * pz[uz[c]];
    cf[1][c] = px[ux[c]] * sy * pz[uz[c]];
    cf[2][c] = px[ux[c]] * py[uy[c]] * sz;
  }
}

// gx[p] = sum_l s_l * sum_ch g[p,l,ch] * d feat[p,l,ch] / d (s_l x)      (what autograd returns for d L / d x)
__global__ void __launch_bounds__(256)
hash_encode_grad_x_kernel(const float* __restrict__ x, int64_t n, const float2* __restrict__ table, const float* __restrict__ scalings,
                          int L, int log2_T, const float* __restrict__ g, float* __restrict__ gx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t mask = (1u << log2_T) - 1u;
  const float px = x[p * 3], py = x[p * 3 + 1], pz = x[p * 3 + 2];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int lev = 0; lev < L; ++lev) {
    const float s = scalings[lev];
    uint32_t idx[8];
    float ox, oy, oz;
    hash_corners(__fmul_rn(px, s), __fmul_rn(py, s), __fmul_rn(pz, s), mask, idx, ox, oy, oz);
    float cf[3][8];
    interp_grad_coefs(ox, oy, oz, cf);
    const float2* tl = table + ((size_t)lev << log2_T);
    const float ga = g[p * 2 * L + 2 * lev], gb = g[p * 2 * L + 2 * lev + 1];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float2 v = __ldg(tl + idx[c]);
      const float t = s * (v.x * ga + v.y * gb);
      a0 = fmaf(cf[0][c], t, a0); a1 = fmaf(cf[1][c], t, a1); a2 = fmaf(cf[2][c], t, a2);
    }
  }
  gx[p * 3] = a0; gx[p * 3 + 1] = a1; gx[p * 3 + 2] = a2;
}

// backward of the above for a cotangent cx [n,3] of gx:
//   d g[p,l,ch]        = s_l * sum_axis cx_axis * d feat[p,l,ch] / d o_axis           (a JVP of the encode along cx)
//   d table[idx_c][ch] += s_l * g[p,l,ch] * sum_axis cx_axis * coef[axis][c]
__global__ void __launch_bounds__(256)
hash_encode_grad_x_bwd_kernel(const float* __restrict__ x, int64_t n, const float2* __restrict__ table, const float* __restrict__ scalings,
                              int L, int log2_T, const float* __restrict__ g, const float* __restrict__ cx, float* __restrict__ d_g,
                              float* __restrict__ d_table) {
  // thread = point walking the levels (see hash_encode_bwd_kernel): the g / d_g rows are read / written once, contiguously
  __shared__ float s_scale[HE_MAX_LEVELS];
  if (threadIdx.x < L) s_scale[threadIdx.x] = scalings[threadIdx.x];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t mask = (1u << log2_T) - 1u;
  const float px = x[p * 3], py = x[p * 3 + 1], pz = x[p * 3 + 2];
  const float c0 = cx[p * 3], c1 = cx[p * 3 + 1], c2 = cx[p * 3 + 2];
  auto level = [&](int lev, float ga, float gb, float& ja, float& jb) {
    const float s = s_scale[lev];
    uint32_t idx[8];
    float ox, oy, oz;
    hash_corners(__fmul_rn(px, s), __fmul_rn(py, s), __fmul_rn(pz, s), mask, idx, ox, oy, oz);
    float cf[3][8];
    interp_grad_coefs(ox, oy, oz, cf);
    const float2* tl = table + ((size_t)lev << log2_T);
    float2* dtl = reinterpret_cast<float2*>(d_table) + ((size_t)lev << log2_T);
    ja = 0.f; jb = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float k = s * (c0 * cf[0][c] + c1 * cf[1][c] + c2 * cf[2][c]);
      const float2 v = __ldg(tl + idx[c]);
      ja = fmaf(k, v.x, ja); jb = fmaf(k, v.y, jb);
      if (d_table) atomicAdd(dtl + idx[c], make_float2(k * ga, k * gb));
    }
  };
  const float* grow = g + p * 2 * L;
  float* drow = d_g ? d_g + p * 2 * L : nullptr;
  if ((L & 1) == 0) {
    for (int lev = 0; lev < L; lev += 2) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(grow) + (lev >> 1));
      float4 j;
      level(lev, gg.x, gg.y, j.x, j.y);
      level(lev + 1, gg.z, gg.w, j.z, j.w);
      if (drow) reinterpret_cast<float4*>(drow)[lev >> 1] = j;
    }
  } else {
    for (int lev = 0; lev < L; ++lev) {
      const float2 gg = __ldg(reinterpret_cast<const float2*>(grow) + lev);
      float2 j;
      level(lev, gg.x, gg.y, j.x, j.y);
      if (drow) reinterpret_cast<float2*>(drow)[lev] = j;
    }
  }
}

// ---- tcnn ("tiny-cuda-nn") grid semantics ---------------------------------------------------------------------------
// The reference's own encodings are tcnn.Encoding modules (sdf_albedo_field.py:117-130, directional_distance_field.py:139-156);
// a trained checkpoint therefore needs tcnn's indexing to be read back (SURVEY A.3 "tcnn differences", restated from memory of
// tiny-cuda-nn grid.h -- unpinned against tcnn itself, pinned by self-consistency tests).  Per level: pos = x * scale + 0.5,
// corners floor / floor + 1, dense index x + y res + z res^2 when res^3 fits the level, else the prime hash, both modulo the
// level's size; linear or smoothstep weights.  meta [L][4] int32 = (float bits of scale, resolution, size, dense);
// the table is our fp32 [L * T, 2] layout filled by tcnn_import.tcnn_params_to_table.
__global__ void __launch_bounds__(256)
hash_encode_tcnn_fwd_kernel(const float* __restrict__ x, int64_t n, const float2* __restrict__ table, const int4* __restrict__ meta,
                            int L, int log2_T, int smoothstep, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (point, level): level-major gathers per warp
  const int lev = blockIdx.y;
  if (i >= n) return;
  const int4 m = meta[lev];
  const float scale = __int_as_float(m.x);
  const uint32_t res = (uint32_t)m.y, size = (uint32_t)m.z;
  const bool dense = m.w != 0;
  float w[3];
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(x[i * 3 + d], scale, 0.5f);
    const float fl = floorf(pos);
    g[d] = (uint32_t)(int)fl;
    const float t = pos - fl;
    w[d] = smoothstep ? t * t * (3.0f - 2.0f * t) : t;
  }
  const float2* tl = table + ((size_t)lev << log2_T);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t gx = g[0] + (c & 1), gy = g[1] + ((c >> 1) & 1), gz = g[2] + ((c >> 2) & 1);
    const float wc = ((c & 1) ? w[0] : 1.f - w[0]) * (((c >> 1) & 1) ? w[1] : 1.f - w[1]) * (((c >> 2) & 1) ? w[2] : 1.f - w[2]);
    const uint32_t idx = (dense ? gx + gy * res + gz * res * res : (gx ^ (gy * 2654435761u) ^ (gz * 805459861u))) % size;
    const float2 f = __ldg(tl + idx);
    acc.x = fmaf(wc, f.x, acc.x);
    acc.y = fmaf(wc, f.y, acc.y);
  }
  reinterpret_cast<float2*>(out)[i * L + lev] = acc;
}

}  // namespace nsk

extern "C" int nsk_hash_encode_grad_x(const float* x, int64_t n, const float* table, const float* scalings, int num_levels,
                                      int log2_T, const float* grad_out, float* grad_x, void* stream) {
  NSK_REQUIRE(num_levels >= 1 && num_levels <= nsk::HE_MAX_LEVELS, "nsk_hash_encode_grad_x: num_levels out of range");
  if (n == 0) return 0;
  NSK_REQUIRE(x && table && scalings && grad_out && grad_x, "nsk_hash_encode_grad_x: null pointer");
  nsk::hash_encode_grad_x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(
      x, n, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, grad_out, grad_x);
  return nsk::check_launch("hash_encode_grad_x_kernel");
}

extern "C" int nsk_hash_encode_grad_x_bwd(const float* x, int64_t n, const float* table, const float* scalings, int num_levels,
                                          int log2_T, const float* grad_out, const float* cot_x, float* d_grad_out,
                                          float* d_table, void* stream) {
  NSK_REQUIRE(num_levels >= 1 && num_levels <= nsk::HE_MAX_LEVELS, "nsk_hash_encode_grad_x_bwd: num_levels out of range");
  if (n == 0) return 0;
  NSK_REQUIRE(x && table && scalings && grad_out && cot_x && (d_grad_out || d_table), "nsk_hash_encode_grad_x_bwd: null pointer");
  const unsigned grid = (unsigned)((n + 255) / 256);
  nsk::hash_encode_grad_x_bwd_kernel<<<grid, 256, 0, nsk::as_stream(stream)>>>(x, n, reinterpret_cast<const float2*>(table), scalings,
                                                                              num_levels, log2_T, grad_out, cot_x, d_grad_out, d_table);
  return nsk::check_launch("hash_encode_grad_x_bwd_kernel");
}

extern "C" int nsk_hash_encode_fwd(const float* x, int64_t n, const float* table, const float* scalings,
                                   int num_levels, int log2_T, float* out, void* stream) {
  NSK_REQUIRE(num_levels >= 1 && num_levels <= nsk::HE_MAX_LEVELS, "nsk_hash_encode_fwd: num_levels out of range");
  NSK_REQUIRE(log2_T >= 1 && log2_T <= 28, "nsk_hash_encode_fwd: log2_T out of range");
  if (n == 0) return 0;
  NSK_REQUIRE(x && table && scalings && out, "nsk_hash_encode_fwd: null pointer");
  int64_t blocks = (n + nsk::HE_WARPS * 32 - 1) / (nsk::HE_WARPS * 32);
  const int64_t cap = 148 * 8 * 4;  // multiple of the SM count; grid-stride beyond that
  if (blocks > cap) blocks = cap;
  nsk::hash_encode_fwd_kernel<<<(unsigned)blocks, nsk::HE_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      x, n, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, out);
  return nsk::check_launch("hash_encode_fwd_kernel");
}

extern "C" int nsk_hash_encode_bwd(const float* x, int64_t n, const float* scalings, int num_levels, int log2_T,
                                   const float* grad_out, float* grad_table, void* stream) {
  NSK_REQUIRE(num_levels >= 1 && num_levels <= nsk::HE_MAX_LEVELS, "nsk_hash_encode_bwd: num_levels out of range");
  if (n == 0) return 0;
  NSK_REQUIRE(x && scalings && grad_out && grad_table, "nsk_hash_encode_bwd: null pointer");
  const unsigned grid = (unsigned)((n + 255) / 256);
  nsk::hash_encode_bwd_kernel<<<grid, 256, 0, nsk::as_stream(stream)>>>(x, n, scalings, num_levels, log2_T, grad_out, grad_table);
  return nsk::check_launch("hash_encode_bwd_kernel");
}

extern "C" int nsk_hash_indices(const float* x, int64_t n, const float* scalings, int num_levels, int log2_T,
                                int64_t* idx, float* offsets, void* stream) {
  if (n == 0) return 0;
  NSK_REQUIRE(x && scalings && idx && offsets, "nsk_hash_indices: null pointer");
  const int64_t total = n * num_levels;
  nsk::hash_indices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(x, n, scalings, num_levels, log2_T, idx, offsets);
  return nsk::check_launch("hash_indices_kernel");
}

extern "C" int nsk_hash_encode_tcnn_fwd(const float* x, int64_t n, const float* table, const int32_t* level_meta, int num_levels, int log2_T,
                                        int smoothstep, float* out, void* stream) {
  NSK_REQUIRE(num_levels >= 1 && num_levels <= nsk::HE_MAX_LEVELS, "nsk_hash_encode_tcnn_fwd: num_levels out of range");
  NSK_REQUIRE(log2_T >= 1 && log2_T <= 28, "nsk_hash_encode_tcnn_fwd: log2_T out of range");
  if (n == 0) return 0;
  NSK_REQUIRE(x && table && level_meta && out, "nsk_hash_encode_tcnn_fwd: null pointer");
  NSK_REQUIRE((reinterpret_cast<uintptr_t>(level_meta) & 15) == 0, "nsk_hash_encode_tcnn_fwd: level_meta must be 16-byte aligned");
  dim3 grid((unsigned)((n + 255) / 256), num_levels);
  nsk::hash_encode_tcnn_fwd_kernel<<<grid, 256, 0, nsk::as_stream(stream)>>>(x, n, reinterpret_cast<const float2*>(table),
                                                                            reinterpret_cast<const int4*>(level_meta), num_levels, log2_T, smoothstep, out);
  return nsk::check_launch("hash_encode_tcnn_fwd_kernel");
}

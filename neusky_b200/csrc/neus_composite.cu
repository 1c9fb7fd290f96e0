// K3: NeuS logistic-CDF alpha + transmittance scan + per-ray composites, one warp per ray.
// Replaces SDFField.get_alpha [NS-mem A.5] (called neusky/fields/sdf_albedo_field.py:266),
// RaySamples.get_weights_and_transmittance_from_alphas (neusky/models/neusky_model.py:565) and the
// accumulation / expected-depth / normal / albedo renderers [NS-mem A.7] (neusky_model.py:591-595,
// 806-813) -- four separate reductions over the same weights in the reference, one pass here.
// Bandwidth-bound: 56*S + 80 algorithmic bytes per ray.
#include "nsk_common.cuh"
#include <float.h>

namespace nsk {

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  // valid for any sign via the int/uint ordering trick
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

constexpr int NC_WARPS = 8;

__global__ void __launch_bounds__(NC_WARPS * 32)
neus_composite_kernel(const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ albedo,
                      const float* __restrict__ ray_dirs, const float* __restrict__ starts, const float* __restrict__ ends,
                      const float* __restrict__ deltas, int64_t R, int S, float inv_s, const float* __restrict__ inv_s_dev, float rho, int training,
                      float* __restrict__ weights, float* __restrict__ wa, float* __restrict__ normals,
                      float* __restrict__ acc_out, float* __restrict__ p2p_raw, float* __restrict__ normal_out,
                      float* __restrict__ albedo_out, float* __restrict__ bgT_out, float* __restrict__ steps_minmax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * NC_WARPS + (threadIdx.x >> 5);
  const int64_t warps_total = (int64_t)gridDim.x * NC_WARPS;
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);   // *_dv entry points: the scalar stays on the device (no host read-back; CUDA-graph capturable)
  float smin = FLT_MAX, smax = -FLT_MAX;
  for (int64_t r = warp_global; r < R; r += warps_total) {
    const float dx = ray_dirs[r * 3], dy = ray_dirs[r * 3 + 1], dz = ray_dirs[r * 3 + 2];
    float T_carry = 1.0f;  // transmittance entering the current chunk of 32 samples
    float a_acc = 0.f, d_acc = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < S;
      const int64_t i = r * S + (ok ? s : S - 1);
      const float sd = sdf[i];
      const float gx = grad[i * 3], gy = grad[i * 3 + 1], gz = grad[i * 3 + 2];
      const float dl = deltas[i];
      const float step = (starts[i] + ends[i]) * 0.5f;
      // NeuS alpha (A.5)
      const float true_cos = dx * gx + dy * gy + dz * gz;
      const float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - rho) + fmaxf(-true_cos, 0.f) * rho);
      const float nxt = sd + iter_cos * dl * 0.5f;
      const float prv = sd - iter_cos * dl * 0.5f;
      const float prev_cdf = sigmoidf_(prv * inv_s);
      const float next_cdf = sigmoidf_(nxt * inv_s);
      float alpha = ((prev_cdf - next_cdf) + 1e-5f) / (prev_cdf + 1e-5f);
      alpha = fminf(fmaxf(alpha, 0.f), 1.f);
      // exclusive product scan of (1 - alpha + 1e-7) across the warp
      float fct = ok ? (1.0f - alpha + 1e-7f) : 1.0f;
      float incl = fct;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = T_carry * excl;
      T_carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (ok) {
        const float w = alpha * T;
        const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);  // F.normalize eps
        const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
        const float ax = albedo[i * 3], ay = albedo[i * 3 + 1], az = albedo[i * 3 + 2];
        weights[i] = w;
        normals[i * 3] = nx; normals[i * 3 + 1] = ny; normals[i * 3 + 2] = nz;
        wa[i * 3] = w * ax; wa[i * 3 + 1] = w * ay; wa[i * 3 + 2] = w * az;
        a_acc += w; d_acc += w * step;
        n0 += w * nx; n1 += w * ny; n2 += w * nz;
        c0 += w * ax; c1 += w * ay; c2 += w * az;
        smin = fminf(smin, step); smax = fmaxf(smax, step);
      }
    }
    a_acc = warp_sum(a_acc); d_acc = warp_sum(d_acc);
    n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (lane == 0) {
      acc_out[r] = a_acc;
      p2p_raw[r] = d_acc / (a_acc + 1e-10f);
      bgT_out[r] = T_carry;
      normal_out[r * 3] = n0; normal_out[r * 3 + 1] = n1; normal_out[r * 3 + 2] = n2;
      float o0 = c0 + (1.0f - a_acc), o1 = c1 + (1.0f - a_acc), o2 = c2 + (1.0f - a_acc);  // white background
      if (!training) { o0 = fminf(fmaxf(o0, 0.f), 1.f); o1 = fminf(fmaxf(o1, 0.f), 1.f); o2 = fminf(fmaxf(o2, 0.f), 1.f); }
      albedo_out[r * 3] = o0; albedo_out[r * 3 + 1] = o1; albedo_out[r * 3 + 2] = o2;
    }
  }
  // batch-global min/max of the sample mid-points (DepthRenderer clips to it, A.7)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  }
  if (lane == 0 && smin <= smax) {
    atomic_min_float(steps_minmax, smin);
    atomic_max_float(steps_minmax + 1, smax);
  }
}

__global__ void neus_finalize_depth_kernel(const float* __restrict__ p2p_raw, const float* __restrict__ dnorm,
                                           const float* __restrict__ mm, int64_t R, float* __restrict__ p2p,
                                           float* __restrict__ depth) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float v = fminf(fmaxf(p2p_raw[r], mm[0]), mm[1]);
  p2p[r] = v;
  depth[r] = v / dnorm[r];
}

__global__ void surface_points_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ p2p,
                                      int64_t R, float radius, float* __restrict__ out) {
  // neusky/models/neusky_model.py:1667-1683 including the element-wise outside-sphere replacement
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float O[3] = {o[r * 3], o[r * 3 + 1], o[r * 3 + 2]};
  const float Dv[3] = {d[r * 3], d[r * 3 + 1], d[r * 3 + 2]};
  float p[3] = {O[0] + Dv[0] * p2p[r], O[1] + Dv[1] * p2p[r], O[2] + Dv[2] * p2p[r]};
  const float nrm = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  if (!(nrm < radius)) {
    float q[3], t;
    sphere_exit(O, Dv, radius, q, t);
    p[0] = q[0] * 0.01f * -Dv[0];
    p[1] = q[1] * 0.01f * -Dv[1];
    p[2] = q[2] * 0.01f * -Dv[2];
  }
  out[r * 3] = p[0]; out[r * 3 + 1] = p[1]; out[r * 3 + 2] = p[2];
}

}  // namespace nsk

static int neus_composite_fwd_impl(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                   const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                   float inv_s, const float* inv_s_dev, float cos_anneal_ratio, int training, float* weights, float* wa,
                                   float* normals, float* acc, float* p2p_raw, float* normal_out, float* albedo_out,
                                   float* bg_T, float* steps_minmax, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1, "nsk_neus_composite_fwd: S must be >= 1");
  NSK_REQUIRE(sdf && grad && albedo && ray_dirs && starts && ends && deltas && weights && wa && normals && acc &&
                  p2p_raw && normal_out && albedo_out && bg_T && steps_minmax,
              "nsk_neus_composite_fwd: null pointer");
  int64_t blocks = (R + nsk::NC_WARPS - 1) / nsk::NC_WARPS;
  const int64_t cap = 148 * 8 * 8;
  if (blocks > cap) blocks = cap;
  nsk::neus_composite_kernel<<<(unsigned)blocks, nsk::NC_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, inv_s, inv_s_dev, cos_anneal_ratio, training, weights, wa, normals,
      acc, p2p_raw, normal_out, albedo_out, bg_T, steps_minmax);
  return nsk::check_launch("neus_composite_kernel");
}

extern "C" int nsk_neus_composite_fwd(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                      const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                      float inv_s, float cos_anneal_ratio, int training, float* weights, float* wa,
                                      float* normals, float* acc, float* p2p_raw, float* normal_out, float* albedo_out,
                                      float* bg_T, float* steps_minmax, void* stream) {
  return neus_composite_fwd_impl(sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, inv_s, nullptr, cos_anneal_ratio, training, weights, wa,
                                 normals, acc, p2p_raw, normal_out, albedo_out, bg_T, steps_minmax, stream);
}

extern "C" int nsk_neus_composite_fwd_dv(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                         const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                         const float* inv_s_dev, float cos_anneal_ratio, int training, float* weights, float* wa,
                                         float* normals, float* acc, float* p2p_raw, float* normal_out, float* albedo_out,
                                         float* bg_T, float* steps_minmax, void* stream) {
  NSK_REQUIRE(inv_s_dev || R == 0, "nsk_neus_composite_fwd_dv: inv_s_dev is NULL");
  return neus_composite_fwd_impl(sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, 0.f, inv_s_dev, cos_anneal_ratio, training, weights, wa,
                                 normals, acc, p2p_raw, normal_out, albedo_out, bg_T, steps_minmax, stream);
}

extern "C" int nsk_neus_finalize_depth(const float* p2p_raw, const float* dnorm, const float* steps_minmax, int64_t R,
                                       float* p2p, float* depth, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(p2p_raw && dnorm && steps_minmax && p2p && depth, "nsk_neus_finalize_depth: null pointer");
  nsk::neus_finalize_depth_kernel<<<(unsigned)((R + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(p2p_raw, dnorm, steps_minmax, R, p2p, depth);
  return nsk::check_launch("neus_finalize_depth_kernel");
}

extern "C" int nsk_surface_points(const float* origins, const float* ray_dirs, const float* p2p, int64_t R, float radius,
                                  float* points, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(origins && ray_dirs && p2p && points, "nsk_surface_points: null pointer");
  nsk::surface_points_kernel<<<(unsigned)((R + 255) / 256), 256, 0, nsk::as_stream(stream)>>>(origins, ray_dirs, p2p, R, radius, points);
  return nsk::check_launch("surface_points_kernel");
}

// =================================================================================================================
// K3 backward: cotangents of every forward output -> d sdf, d gradient, d albedo (per sample) and d inv_s (scalar).
// Same algebra that torch autograd derives for SDFField.get_alpha + get_weights_and_transmittance_from_alphas + the
// renderers; one warp per ray, alpha / transmittance recomputed (nothing but the inputs is saved).
//   w_s = alpha_s T_s,  T_{s+1} = T_s (1 - alpha_s + 1e-7):
//   dL/dalpha_s = gw_s T_s - (sum_{k>s} gw_k w_k + g_bgT T_S) / (1 - alpha_s + 1e-7)        (gw = total cotangent of w)
// =================================================================================================================
namespace nsk {

constexpr int NCB_MAX_CHUNKS = 32;   // S <= 1024

__global__ void __launch_bounds__(NC_WARPS * 32)
neus_composite_bwd_kernel(const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ albedo,
                          const float* __restrict__ ray_dirs, const float* __restrict__ starts, const float* __restrict__ ends,
                          const float* __restrict__ deltas, int64_t R, int S, float inv_s, const float* __restrict__ inv_s_dev, float rho,
                          const float* __restrict__ g_weights, const float* __restrict__ g_wa, const float* __restrict__ g_normals,
                          const float* __restrict__ g_acc, const float* __restrict__ g_p2p_raw, const float* __restrict__ g_normal_out,
                          const float* __restrict__ g_albedo_out, const float* __restrict__ g_bgT,
                          float* __restrict__ d_sdf, float* __restrict__ d_grad, float* __restrict__ d_albedo, float* __restrict__ d_inv_s) {
  __shared__ float s_Tin[NC_WARPS][NCB_MAX_CHUNKS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * NC_WARPS + wib;
  const int64_t warps_total = (int64_t)gridDim.x * NC_WARPS;
  const int nchunks = (S + 31) / 32;
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);
  float dinv_acc = 0.f;

  for (int64_t r = warp_global; r < R; r += warps_total) {
    const float dx = ray_dirs[r * 3], dy = ray_dirs[r * 3 + 1], dz = ray_dirs[r * 3 + 2];
    // alpha of sample i (recomputed); also returns the pieces the chain rule needs
    auto alpha_of = [&](int64_t i, float& pc, float& nc, float& prv, float& nxt, float& tcos, bool& clipped) {
      const float sd = sdf[i], dl = deltas[i];
      tcos = dx * grad[i * 3] + dy * grad[i * 3 + 1] + dz * grad[i * 3 + 2];
      const float ic = -(fmaxf(-tcos * 0.5f + 0.5f, 0.f) * (1.0f - rho) + fmaxf(-tcos, 0.f) * rho);
      nxt = sd + ic * dl * 0.5f;
      prv = sd - ic * dl * 0.5f;
      pc = sigmoidf_(prv * inv_s);
      nc = sigmoidf_(nxt * inv_s);
      const float a = ((pc - nc) + 1e-5f) / (pc + 1e-5f);
      clipped = !(a >= 0.f && a <= 1.f);
      return fminf(fmaxf(a, 0.f), 1.f);
    };
    // ---- sweep 1: transmittance entering each 32-sample chunk, and the per-ray sums ----------------------------
    float T_carry = 1.0f, acc = 0.f, dsum = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const int64_t i = r * S + (ok ? s : S - 1);
      float pc, nc, prv, nxt, tc; bool cl;
      const float alpha = alpha_of(i, pc, nc, prv, nxt, tc, cl);
      float incl = ok ? (1.0f - alpha + 1e-7f) : 1.0f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= up; }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      if (lane == 0) s_Tin[wib][c] = T_carry;
      const float w = ok ? alpha * T_carry * excl : 0.f;
      acc += w;
      dsum += w * (starts[i] + ends[i]) * 0.5f;
      T_carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    acc = warp_sum(acc); dsum = warp_sum(dsum);
    __syncwarp();
    const float T_end = T_carry;
    const float gacc = g_acc ? g_acc[r] : 0.f, gp2p = g_p2p_raw ? g_p2p_raw[r] : 0.f, gbg = g_bgT ? g_bgT[r] : 0.f;
    float gno[3] = {0.f, 0.f, 0.f}, gao[3] = {0.f, 0.f, 0.f};
    if (g_normal_out) { gno[0] = g_normal_out[r * 3]; gno[1] = g_normal_out[r * 3 + 1]; gno[2] = g_normal_out[r * 3 + 2]; }
    if (g_albedo_out) { gao[0] = g_albedo_out[r * 3]; gao[1] = g_albedo_out[r * 3 + 1]; gao[2] = g_albedo_out[r * 3 + 2]; }
    const float den = acc + 1e-10f;
    // ---- sweep 2: chunks in reverse, suffix sums of gw_k w_k ---------------------------------------------------------
    float suffix = gbg * T_end;
    for (int c = nchunks - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const int64_t i = r * S + (ok ? s : S - 1);
      float pc, nc, prv, nxt, tc; bool cl;
      const float alpha = alpha_of(i, pc, nc, prv, nxt, tc, cl);
      const float fct = ok ? (1.0f - alpha + 1e-7f) : 1.0f;
      float incl = fct;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= up; }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = s_Tin[wib][c] * excl;
      const float w = ok ? alpha * T : 0.f;
      const float gx = grad[i * 3], gy = grad[i * 3 + 1], gz = grad[i * 3 + 2];
      const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
      const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
      const float ax = albedo[i * 3], ay = albedo[i * 3 + 1], az = albedo[i * 3 + 2];
      const float mid = (starts[i] + ends[i]) * 0.5f;
      float gwa0 = 0.f, gwa1 = 0.f, gwa2 = 0.f;
      if (g_wa) { gwa0 = g_wa[i * 3]; gwa1 = g_wa[i * 3 + 1]; gwa2 = g_wa[i * 3 + 2]; }
      // total cotangent of w_s
      float gw = (g_weights ? g_weights[i] : 0.f) + gacc + gp2p * (mid / den - dsum / (den * den)) + gno[0] * nx + gno[1] * ny + gno[2] * nz +
                 gao[0] * (ax - 1.0f) + gao[1] * (ay - 1.0f) + gao[2] * (az - 1.0f) + gwa0 * ax + gwa1 * ay + gwa2 * az;
      if (!ok) gw = 0.f;
      // suffix over k > s within the chunk (reverse exclusive scan of gw_k w_k) + everything after the chunk
      const float v = gw * w;
      float rincl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float dn = __shfl_down_sync(0xffffffffu, rincl, o); if (lane + o < 32) rincl += dn; }
      float rexcl = __shfl_down_sync(0xffffffffu, rincl, 1);
      if (lane == 31) rexcl = 0.f;
      const float after = rexcl + suffix;
      suffix += __shfl_sync(0xffffffffu, rincl, 0);
      if (ok) {
        float dalpha = gw * T - after / fct;
        if (cl) dalpha = 0.f;                                     // clip(0,1) passes no gradient outside the interval
        const float dpc = dalpha * nc / ((pc + 1e-5f) * (pc + 1e-5f));
        const float dnc = -dalpha / (pc + 1e-5f);
        const float dprv = dpc * pc * (1.0f - pc) * inv_s, dnxt = dnc * nc * (1.0f - nc) * inv_s;
        dinv_acc += dpc * pc * (1.0f - pc) * prv + dnc * nc * (1.0f - nc) * nxt;
        d_sdf[i] = dprv + dnxt;
        const float dic = (dnxt - dprv) * deltas[i] * 0.5f;
        const float dtc = dic * (((-tc * 0.5f + 0.5f) > 0.f ? (1.0f - rho) * 0.5f : 0.f) + ((-tc) > 0.f ? rho : 0.f));
        // normals_s = g/|g|: cotangent from the rendered normal and from the shading pass
        float dn0 = w * gno[0], dn1 = w * gno[1], dn2 = w * gno[2];
        if (g_normals) { dn0 += g_normals[i * 3]; dn1 += g_normals[i * 3 + 1]; dn2 += g_normals[i * 3 + 2]; }
        const float dot = dn0 * nx + dn1 * ny + dn2 * nz;
        d_grad[i * 3] = dtc * dx + (dn0 - dot * nx) / gn;
        d_grad[i * 3 + 1] = dtc * dy + (dn1 - dot * ny) / gn;
        d_grad[i * 3 + 2] = dtc * dz + (dn2 - dot * nz) / gn;
        d_albedo[i * 3] = w * (gao[0] + gwa0);
        d_albedo[i * 3 + 1] = w * (gao[1] + gwa1);
        d_albedo[i * 3 + 2] = w * (gao[2] + gwa2);
      }
    }
    __syncwarp();
  }
  dinv_acc = warp_sum(dinv_acc);
  if (lane == 0 && dinv_acc != 0.f) atomicAdd(d_inv_s, dinv_acc);
}

}  // namespace nsk

static int neus_composite_bwd_impl(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                   const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                   float inv_s, const float* inv_s_dev, float cos_anneal_ratio, const float* g_weights, const float* g_wa,
                                   const float* g_normals, const float* g_acc, const float* g_p2p_raw,
                                   const float* g_normal_out, const float* g_albedo_out, const float* g_bg_T,
                                   float* d_sdf, float* d_grad, float* d_albedo, float* d_inv_s, void* stream) {
  if (R == 0) return 0;
  NSK_REQUIRE(S >= 1 && S <= 32 * nsk::NCB_MAX_CHUNKS, "nsk_neus_composite_bwd: S must be in [1, 1024]");
  NSK_REQUIRE(sdf && grad && albedo && ray_dirs && starts && ends && deltas && d_sdf && d_grad && d_albedo && d_inv_s,
              "nsk_neus_composite_bwd: null pointer");
  int64_t blocks = (R + nsk::NC_WARPS - 1) / nsk::NC_WARPS;
  const int64_t cap = 148 * 8 * 8;
  if (blocks > cap) blocks = cap;
  nsk::neus_composite_bwd_kernel<<<(unsigned)blocks, nsk::NC_WARPS * 32, 0, nsk::as_stream(stream)>>>(
      sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, inv_s, inv_s_dev, cos_anneal_ratio, g_weights, g_wa, g_normals, g_acc,
      g_p2p_raw, g_normal_out, g_albedo_out, g_bg_T, d_sdf, d_grad, d_albedo, d_inv_s);
  return nsk::check_launch("neus_composite_bwd_kernel");
}

extern "C" int nsk_neus_composite_bwd(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                      const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                      float inv_s, float cos_anneal_ratio, const float* g_weights, const float* g_wa,
                                      const float* g_normals, const float* g_acc, const float* g_p2p_raw,
                                      const float* g_normal_out, const float* g_albedo_out, const float* g_bg_T,
                                      float* d_sdf, float* d_grad, float* d_albedo, float* d_inv_s, void* stream) {
  return neus_composite_bwd_impl(sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, inv_s, nullptr, cos_anneal_ratio, g_weights, g_wa, g_normals,
                                 g_acc, g_p2p_raw, g_normal_out, g_albedo_out, g_bg_T, d_sdf, d_grad, d_albedo, d_inv_s, stream);
}

extern "C" int nsk_neus_composite_bwd_dv(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                                         const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                                         const float* inv_s_dev, float cos_anneal_ratio, const float* g_weights, const float* g_wa,
                                         const float* g_normals, const float* g_acc, const float* g_p2p_raw,
                                         const float* g_normal_out, const float* g_albedo_out, const float* g_bg_T,
                                         float* d_sdf, float* d_grad, float* d_albedo, float* d_inv_s, void* stream) {
  NSK_REQUIRE(inv_s_dev || R == 0, "nsk_neus_composite_bwd_dv: inv_s_dev is NULL");
  return neus_composite_bwd_impl(sdf, grad, albedo, ray_dirs, starts, ends, deltas, R, S, 0.f, inv_s_dev, cos_anneal_ratio, g_weights, g_wa, g_normals,
                                 g_acc, g_p2p_raw, g_normal_out, g_albedo_out, g_bg_T, d_sdf, d_grad, d_albedo, d_inv_s, stream);
}

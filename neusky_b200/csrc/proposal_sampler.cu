// Proposal-network sampler: the step directly in front of the render-and-shade path.
// replaces nerfstudio's ProposalNetworkSampler.generate_ray_samples as NeuSky calls it at
// neusky/models/neusky_model.py:561 (UniformSampler -> HashMLPDensityField -> RaySamples.get_weights -> PDFSampler, twice)
// [NS-mem; restated in oracle/sampler_oracle.py].
//
//   P1  proposal_density_fwd / bwd   one thread per sample: position from (ray, bin) -> L-inf contraction -> (p+2)/4 ->
//                                    selector -> hash encode (5 levels x 8 corners x 8 B from a 5 MB, L2-resident table) ->
//                                    16-wide MLP -> trunc_exp.  HBM traffic 4 B in (one bin edge) + 4 B out per sample; the
//                                    algorithmic figure of SURVEY 8d (372 B/point incl. gathers) is what the roofline quotes.
//   P2  pdf_resample                 one warp per ray: density -> weights (exclusive scan) -> annealed, padded pdf -> cdf
//                                    (inclusive scan) -> searchsorted of the N+1 sample positions -> new bins.
//                                    Scans accumulate in fp64 like torch's CPU cumsum.  With histogram_padding = 0.01 every
//                                    pdf entry is >= 2^-9 and every prefix < 2, so all fp64 partial sums are EXACT and the
//                                    parallel scan is bit-identical to the sequential one: sample placement is bit-exact
//                                    given the weights.
//   P3  interlevel loss              one warp per ray (lossfun_outer of the proposal histogram against the fine one).
#include "nsk_common.cuh"

namespace nsk {

constexpr int PS_MAX_LEVELS = 8;
constexpr int PS_HID = 16;
constexpr int PS_MAX_S = 512;   // samples per ray a P2 warp can hold (16 per lane)

// UniformSampler: spacing_fn = identity -> x*far + (1-x)*near, one rounding per op (no FMA contraction)
__device__ __forceinline__ float sp2e(float b, float nr, float fr) {
  return __fadd_rn(__fmul_rn(b, fr), __fmul_rn(__fsub_rn(1.0f, b), nr));
}

// position of sample s of ray r (Frustums.get_positions: o + d * (start+end)/2), contracted and mapped to [0,1]^3
__device__ __forceinline__ bool proposal_position(const float* __restrict__ origins, const float* __restrict__ dirs,
                                                  const float* __restrict__ near, const float* __restrict__ far,
                                                  const float* __restrict__ bins, int64_t r, int s, int S, float x[3]) {
  float pos[3];
  if (dirs == nullptr) {   // positions mode: `origins` holds [R*S,3] world positions
    const int64_t i = r * S + s;
    pos[0] = origins[i * 3]; pos[1] = origins[i * 3 + 1]; pos[2] = origins[i * 3 + 2];
  } else {
    const float nr = near[r], fr = far[r];
    const float e0 = sp2e(bins[r * (S + 1) + s], nr, fr), e1 = sp2e(bins[r * (S + 1) + s + 1], nr, fr);
    const float mid = __fmul_rn(__fadd_rn(e0, e1), 0.5f);
#pragma unroll
    for (int j = 0; j < 3; ++j) pos[j] = __fadd_rn(origins[r * 3 + j], __fmul_rn(dirs[r * 3 + j], mid));
  }
  const float mag = fmaxf(fabsf(pos[0]), fmaxf(fabsf(pos[1]), fabsf(pos[2])));
  bool sel = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float c = pos[j];
    if (!(mag < 1.0f)) c = __fmul_rn(__fsub_rn(2.0f, __fdiv_rn(1.0f, mag)), __fdiv_rn(pos[j], mag));
    c = __fmul_rn(__fadd_rn(c, 2.0f), 0.25f);
    sel = sel && (c > 0.0f) && (c < 1.0f);
    x[j] = c;
  }
  if (!sel) x[0] = x[1] = x[2] = 0.0f;
  return sel;
}

// packed MLP blob: W0 [2L][16] (input-major), b0 [16], W1 [16], b1 [1]
__device__ __forceinline__ int mlp_floats(int L) { return 2 * L * PS_HID + PS_HID + PS_HID + 1; }

__global__ void __launch_bounds__(256)
proposal_density_fwd_kernel(const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ near,
                            const float* __restrict__ far, const float* __restrict__ bins, int64_t R, int S,
                            const float2* __restrict__ table, const float* __restrict__ scalings, int L, int log2_T,
                            const float* __restrict__ mlp, float* __restrict__ density) {
  __shared__ float s_mlp[2 * PS_MAX_LEVELS * PS_HID + 2 * PS_HID + 1];
  __shared__ float s_scale[PS_MAX_LEVELS];
  for (int i = threadIdx.x; i < mlp_floats(L); i += blockDim.x) s_mlp[i] = mlp[i];
  if (threadIdx.x < L) s_scale[threadIdx.x] = scalings[threadIdx.x];
  __syncthreads();
  const uint32_t mask = (1u << log2_T) - 1u;
  const int64_t total = R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / S;
    const int s = (int)(i - r * S);
    float x[3];
    const bool sel = proposal_position(origins, dirs, near, far, bins, r, s, S, x);
    float h[PS_HID];
#pragma unroll
    for (int j = 0; j < PS_HID; ++j) h[j] = s_mlp[2 * L * PS_HID + j];
    for (int l = 0; l < L; ++l) {
      const float sc = s_scale[l];
      uint32_t idx[8];
      float ox, oy, oz;
      hash_corners(__fmul_rn(x[0], sc), __fmul_rn(x[1], sc), __fmul_rn(x[2], sc), mask, idx, ox, oy, oz);
      const float2* tl = table + ((size_t)l << log2_T);
      float2 f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
      const float2 v = hash_interp(f, ox, oy, oz);
      const float* w0 = s_mlp + (2 * l) * PS_HID;
#pragma unroll
      for (int j = 0; j < PS_HID; ++j) h[j] = fmaf(v.x, w0[j], fmaf(v.y, w0[PS_HID + j], h[j]));
    }
    float raw = s_mlp[2 * L * PS_HID + 2 * PS_HID];
    const float* w1 = s_mlp + 2 * L * PS_HID + PS_HID;
#pragma unroll
    for (int j = 0; j < PS_HID; ++j) raw = fmaf(fmaxf(h[j], 0.f), w1[j], raw);
    density[i] = sel ? expf(raw) : 0.0f;
  }
}

// Backward of P1 for the proposal-network parameters ("proposal_networks" optimizer group, neusky_model.py:379-398):
// g_density [R,S] -> d_table [L*T,2] (atomics), d_mlp [mlp_floats] (block-reduced, then atomics).  trunc_exp backward:
// d raw = g * exp(min(raw, 15)) [NS-mem].
__global__ void __launch_bounds__(256)
proposal_density_bwd_kernel(const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ near,
                            const float* __restrict__ far, const float* __restrict__ bins, int64_t R, int S,
                            const float2* __restrict__ table, const float* __restrict__ scalings, int L, int log2_T,
                            const float* __restrict__ mlp, const float* __restrict__ g_density,
                            float* __restrict__ d_table, float* __restrict__ d_mlp) {
  __shared__ float s_mlp[2 * PS_MAX_LEVELS * PS_HID + 2 * PS_HID + 1];
  __shared__ float s_dmlp[2 * PS_MAX_LEVELS * PS_HID + 2 * PS_HID + 1];
  __shared__ float s_scale[PS_MAX_LEVELS];
  const int nm = mlp_floats(L);
  for (int i = threadIdx.x; i < nm; i += blockDim.x) { s_mlp[i] = mlp[i]; s_dmlp[i] = 0.f; }
  if (threadIdx.x < L) s_scale[threadIdx.x] = scalings[threadIdx.x];
  __syncthreads();
  const uint32_t mask = (1u << log2_T) - 1u;
  const int64_t total = R * S;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (total + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < total;
    float x[3] = {0.f, 0.f, 0.f};
    bool sel = false;
    float g = 0.f;
    if (live) {
      const int64_t r = i / S;
      sel = proposal_position(origins, dirs, near, far, bins, r, (int)(i - r * S), S, x);
      g = g_density[i];
    }
    // recompute the forward
    float feat[2 * PS_MAX_LEVELS];
    float h[PS_HID];
#pragma unroll
    for (int j = 0; j < PS_HID; ++j) h[j] = s_mlp[2 * L * PS_HID + j];
    for (int l = 0; l < L; ++l) {
      const float sc = s_scale[l];
      uint32_t idx[8];
      float ox, oy, oz;
      hash_corners(__fmul_rn(x[0], sc), __fmul_rn(x[1], sc), __fmul_rn(x[2], sc), mask, idx, ox, oy, oz);
      const float2* tl = table + ((size_t)l << log2_T);
      float2 f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
      const float2 v = hash_interp(f, ox, oy, oz);
      feat[2 * l] = v.x; feat[2 * l + 1] = v.y;
      const float* w0 = s_mlp + (2 * l) * PS_HID;
#pragma unroll
      for (int j = 0; j < PS_HID; ++j) h[j] = fmaf(v.x, w0[j], fmaf(v.y, w0[PS_HID + j], h[j]));
    }
    float raw = s_mlp[2 * L * PS_HID + 2 * PS_HID];
    const float* w1 = s_mlp + 2 * L * PS_HID + PS_HID;
#pragma unroll
    for (int j = 0; j < PS_HID; ++j) raw = fmaf(fmaxf(h[j], 0.f), w1[j], raw);
    const float d_raw = (live && sel) ? g * expf(fminf(raw, 15.0f)) : 0.f;
    // ---- MLP parameter gradients: warp-reduce, one shared atomic per warp and parameter
    float dh[PS_HID];
#pragma unroll
    for (int j = 0; j < PS_HID; ++j) {
      const float a = fmaxf(h[j], 0.f);
      dh[j] = h[j] > 0.f ? d_raw * w1[j] : 0.f;
      const float gw1 = warp_sum(d_raw * a);
      const float gb0 = warp_sum(dh[j]);
      if (lane == 0) {
        atomicAdd(&s_dmlp[2 * L * PS_HID + PS_HID + j], gw1);
        atomicAdd(&s_dmlp[2 * L * PS_HID + j], gb0);
      }
    }
    {
      const float gb1 = warp_sum(d_raw);
      if (lane == 0) atomicAdd(&s_dmlp[2 * L * PS_HID + 2 * PS_HID], gb1);
    }
    for (int k = 0; k < 2 * L; ++k) {
#pragma unroll
      for (int j = 0; j < PS_HID; ++j) {
        const float gw0 = warp_sum(feat[k] * dh[j]);
        if (lane == 0) atomicAdd(&s_dmlp[k * PS_HID + j], gw0);
      }
    }
    // ---- table gradient through the trilinear weights
    if (d_raw != 0.f) {
      const int ux[8] = {0, 0, 1, 1, 0, 0, 1, 1}, uy[8] = {0, 1, 1, 0, 0, 1, 1, 0}, uz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
      for (int l = 0; l < L; ++l) {
        const float* w0 = s_mlp + (2 * l) * PS_HID;
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int j = 0; j < PS_HID; ++j) { gx = fmaf(dh[j], w0[j], gx); gy = fmaf(dh[j], w0[PS_HID + j], gy); }
        const float sc = s_scale[l];
        uint32_t idx[8];
        float ox, oy, oz;
        hash_corners(__fmul_rn(x[0], sc), __fmul_rn(x[1], sc), __fmul_rn(x[2], sc), mask, idx, ox, oy, oz);
        const float wx[2] = {ox, 1.f - ox}, wy[2] = {oy, 1.f - oy}, wz[2] = {oz, 1.f - oz};
        float2* tl = reinterpret_cast<float2*>(d_table) + ((size_t)l << log2_T);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float w = wx[ux[c]] * wy[uy[c]] * wz[uz[c]];
          atomicAdd(tl + idx[c], make_float2(w * gx, w * gy));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nm; i += blockDim.x)
    if (s_dmlp[i] != 0.f) atomicAdd(d_mlp + i, s_dmlp[i]);
}

// ---- warp scans in fp64 -----------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_excl_scan(double v, double& total) {
  const int lane = threadIdx.x & 31;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// SpacedSampler bins [NS-mem A.6]: base = linspace(0,1,S+1) (host, CPU rounding); training with single jitter:
// bins = lower + (upper-lower)*t_rand, lower/upper = the bin-centre brackets.
__global__ void uniform_bins_kernel(const float* __restrict__ base, const float* __restrict__ jitter, int64_t R, int S,
                                    float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * (S + 1)) return;
  const int64_t r = i / (S + 1);
  const int k = (int)(i - r * (S + 1));
  float b = base[k];
  if (jitter != nullptr) {
    const float upper = k < S ? __fmul_rn(__fadd_rn(base[k + 1], base[k]), 0.5f) : base[S];
    const float lower = k > 0 ? __fmul_rn(__fadd_rn(base[k], base[k - 1]), 0.5f) : base[0];
    b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), jitter[r]));
  }
  out[i] = b;
}

// P2: one warp per ray.
constexpr int PR_WARPS = 4;
__global__ void __launch_bounds__(PR_WARPS * 32)
pdf_resample_kernel(const float* __restrict__ bins, const float* __restrict__ density, const float* __restrict__ weights_in,
                    const float* __restrict__ near, const float* __restrict__ far, int64_t R, int S, int N, float anneal,
                    float hist_pad, float eps, const float* __restrict__ u_base, float u_half,
                    const float* __restrict__ jitter, float* __restrict__ weights_out, float* __restrict__ new_bins,
                    float* __restrict__ new_euclid) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_bins = smem + (size_t)warp * 2 * (S + 1);
  float* s_cdf = s_bins + (S + 1);
  const int C = (S + 31) / 32;   // elements per lane (contiguous chunk)
  for (int64_t r = (int64_t)blockIdx.x * PR_WARPS + warp; r < R; r += (int64_t)gridDim.x * PR_WARPS) {
    const float nr = near[r], fr = far[r];
    for (int i = lane; i <= S; i += 32) s_bins[i] = bins[r * (S + 1) + i];
    __syncwarp();
    const int i0 = lane * C;
    float w[PS_MAX_S / 32];
    if (weights_in != nullptr) {
#pragma unroll
      for (int k = 0; k < PS_MAX_S / 32; ++k) w[k] = (k < C && i0 + k < S) ? weights_in[r * S + i0 + k] : 0.f;
    } else {
      // ---- RaySamples.get_weights: dd = delta*density; alpha = 1-exp(-dd); T = exp(-cumsum([0, dd[:-1]])) ----
      float dd[PS_MAX_S / 32];
      double loc = 0.0;
#pragma unroll
      for (int k = 0; k < PS_MAX_S / 32; ++k) {
        dd[k] = 0.f;
        if (k < C && i0 + k < S) {
          const float delta = __fsub_rn(sp2e(s_bins[i0 + k + 1], nr, fr), sp2e(s_bins[i0 + k], nr, fr));
          dd[k] = __fmul_rn(delta, density[r * S + i0 + k]);
          loc += (double)dd[k];
        }
      }
      double tot;
      double run = warp_excl_scan(loc, tot);
#pragma unroll
      for (int k = 0; k < PS_MAX_S / 32; ++k) {
        w[k] = 0.f;
        if (k < C && i0 + k < S) {
          const float alpha = __fsub_rn(1.0f, expf(-dd[k]));
          const float T = expf(-(float)run);
          float v = __fmul_rn(alpha, T);
          if (isnan(v)) v = 0.f;
          else if (isinf(v)) v = v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
          w[k] = v;
          run += (double)dd[k];
          if (weights_out != nullptr) weights_out[r * S + i0 + k] = v;
        }
      }
    }
    // ---- PDFSampler: anneal, pad, normalise ----
    double loc = 0.0;
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      if (k < C && i0 + k < S) {
        float a = w[k];
        if (anneal != 1.0f) a = powf(a, anneal);
        w[k] = __fadd_rn(a, hist_pad);
        loc += (double)w[k];
      }
    }
    double tot;
    warp_excl_scan(loc, tot);
    float wsum = (float)tot;
    const float pad = fmaxf(__fsub_rn(eps, wsum), 0.0f);
    const float padw = __fdiv_rn(pad, (float)S);
    wsum = __fadd_rn(wsum, pad);
    loc = 0.0;
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      if (k < C && i0 + k < S) {
        w[k] = __fdiv_rn(__fadd_rn(w[k], padw), wsum);   // pdf
        loc += (double)w[k];
      }
    }
    double run = warp_excl_scan(loc, tot);
    if (lane == 0) s_cdf[0] = 0.f;
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      if (k < C && i0 + k < S) {
        run += (double)w[k];
        s_cdf[i0 + k + 1] = fminf(1.0f, (float)run);
      }
    }
    __syncwarp();
    // ---- inverse-CDF sampling of the N+1 new bin edges ----
    const float shift = jitter != nullptr ? __fdiv_rn(jitter[r], (float)(N + 1)) : u_half;
    for (int j = lane; j <= N; j += 32) {
      const float u = __fadd_rn(u_base[j], shift);
      int lo = 0, hi = S + 1;                       // searchsorted(cdf, u, right): first index with cdf > u
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int below = min(max(lo - 1, 0), S), above = min(max(lo, 0), S);
      const float c0 = s_cdf[below], c1 = s_cdf[above], b0 = s_bins[below], b1 = s_bins[above];
      float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
      if (isnan(t)) t = 0.f;
      t = fminf(fmaxf(t, 0.f), 1.f);
      const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      new_bins[r * (N + 1) + j] = nbv;
      if (new_euclid != nullptr) new_euclid[r * (N + 1) + j] = sp2e(nbv, nr, fr);
    }
    __syncwarp();
  }
}

// Backward of RaySamples.get_weights wrt the density (the only path by which the interlevel loss reaches the proposal
// networks): w_i = (1-exp(-dd_i)) exp(-c_i), c_i = sum_{k<i} dd_k  ->  g_dd_i = g_w_i exp(-dd_i) T_i - sum_{j>i} g_w_j w_j.
__global__ void __launch_bounds__(PR_WARPS * 32)
density_weights_bwd_kernel(const float* __restrict__ bins, const float* __restrict__ density, const float* __restrict__ near,
                           const float* __restrict__ far, int64_t R, int S, const float* __restrict__ g_w,
                           float* __restrict__ g_density) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = (S + 31) / 32;
  for (int64_t r = (int64_t)blockIdx.x * PR_WARPS + warp; r < R; r += (int64_t)gridDim.x * PR_WARPS) {
    const float nr = near[r], fr = far[r];
    const int i0 = lane * C;
    float dd[PS_MAX_S / 32], dl[PS_MAX_S / 32], gw[PS_MAX_S / 32];
    double loc = 0.0;
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      dd[k] = dl[k] = gw[k] = 0.f;
      if (k < C && i0 + k < S) {
        dl[k] = __fsub_rn(sp2e(bins[r * (S + 1) + i0 + k + 1], nr, fr), sp2e(bins[r * (S + 1) + i0 + k], nr, fr));
        dd[k] = dl[k] * density[r * S + i0 + k];
        gw[k] = g_w[r * S + i0 + k];
        loc += (double)dd[k];
      }
    }
    double tot;
    double run = warp_excl_scan(loc, tot);
    float wv[PS_MAX_S / 32], T[PS_MAX_S / 32];
    double locg = 0.0;
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      wv[k] = T[k] = 0.f;
      if (k < C && i0 + k < S) {
        T[k] = expf(-(float)run);
        wv[k] = (1.0f - expf(-dd[k])) * T[k];
        run += (double)dd[k];
        locg += (double)(gw[k] * wv[k]);
      }
    }
    double totg;
    const double before = warp_excl_scan(locg, totg);
    double suffix = totg - before;   // sum over this lane's chunk and everything after it
#pragma unroll
    for (int k = 0; k < PS_MAX_S / 32; ++k) {
      if (k < C && i0 + k < S) {
        suffix -= (double)(gw[k] * wv[k]);           // now: sum_{j>i}
        const float gdd = gw[k] * expf(-dd[k]) * T[k] - (float)suffix;
        g_density[r * S + i0 + k] = gdd * dl[k];
      }
    }
  }
}

// P3: nerfstudio interlevel_loss term of one proposal level [NS-mem]: lossfun_outer(c, w, cp, wp) per fine interval,
//   w_outer_i = cy[hi_i + 1] - cy[lo_i], cy = [0, cumsum(wp)], lo = clamp(searchsorted(cp[:-1], c_i, right) - 1),
//   hi = clamp(searchsorted(cp[1:], c_{i+1}, right)), loss_i = relu(w_i - w_outer_i)^2 / (w_i + eps).
// Writes the per-ray sum of loss_i (host takes the mean over R*Sf) and, if g_wp != NULL, d(sum over rays)/d wp.
__global__ void __launch_bounds__(PR_WARPS * 32)
interlevel_loss_kernel(const float* __restrict__ c, const float* __restrict__ w, int Sf, const float* __restrict__ cp,
                       const float* __restrict__ wp, int Sp, int64_t R, float* __restrict__ loss_ray, float* __restrict__ g_wp) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_cp = smem + (size_t)warp * (3 * (Sp + 1));
  float* s_cy = s_cp + (Sp + 1);
  float* s_g = s_cy + (Sp + 1);     // d loss / d cy
  const float EPS = 1.1920928955078125e-07f;
  const int C = (Sp + 31) / 32;
  for (int64_t r = (int64_t)blockIdx.x * PR_WARPS + warp; r < R; r += (int64_t)gridDim.x * PR_WARPS) {
    for (int i = lane; i <= Sp; i += 32) { s_cp[i] = cp[r * (Sp + 1) + i]; s_g[i] = 0.f; }
    const int i0 = lane * C;
    double loc = 0.0;
    for (int k = 0; k < C; ++k) if (i0 + k < Sp) loc += (double)wp[r * Sp + i0 + k];
    double tot;
    double run = warp_excl_scan(loc, tot);
    if (lane == 0) s_cy[0] = 0.f;
    for (int k = 0; k < C; ++k) if (i0 + k < Sp) { run += (double)wp[r * Sp + i0 + k]; s_cy[i0 + k + 1] = (float)run; }
    __syncwarp();
    float acc = 0.f;
    for (int i = lane; i < Sf; i += 32) {
      const float t0 = c[r * (Sf + 1) + i], t1 = c[r * (Sf + 1) + i + 1], wi = w[r * Sf + i];
      int lo = 0, hi = Sp;                            // over cp[0..Sp-1]
      while (lo < hi) { const int m = (lo + hi) >> 1; if (s_cp[m] <= t0) lo = m + 1; else hi = m; }
      const int idx_lo = min(max(lo - 1, 0), Sp - 1);
      lo = 0; hi = Sp;                                // over cp[1..Sp]
      while (lo < hi) { const int m = (lo + hi) >> 1; if (s_cp[m + 1] <= t1) lo = m + 1; else hi = m; }
      const int idx_hi = min(max(lo, 0), Sp - 1);
      const float outer = s_cy[idx_hi + 1] - s_cy[idx_lo];
      const float d = fmaxf(wi - outer, 0.f);
      acc += d * d / (wi + EPS);
      if (g_wp != nullptr && d > 0.f) {
        const float gouter = -2.f * d / (wi + EPS);
        atomicAdd(&s_g[idx_hi + 1], gouter);
        atomicAdd(&s_g[idx_lo], -gouter);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) loss_ray[r] = acc;
    __syncwarp();
    if (g_wp != nullptr) {
      // cy[k] = sum_{j<k} wp_j  ->  d/d wp_j = sum_{k>j} g_cy[k]: suffix sums
      double lg = 0.0;
      for (int k = 0; k < C; ++k) if (i0 + k < Sp) lg += (double)s_g[i0 + k + 1];
      double tg;
      const double before = warp_excl_scan(lg, tg);
      double suffix = tg - before;
      for (int k = 0; k < C; ++k) if (i0 + k < Sp) { g_wp[r * Sp + i0 + k] = (float)suffix; suffix -= (double)s_g[i0 + k + 1]; }
    }
    __syncwarp();
  }
}

static int grid_for(int64_t items, int per_block, int cap = 148 * 16) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

}  // namespace nsk

using namespace nsk;

extern "C" int64_t nsk_proposal_mlp_floats(int num_levels, int hidden) {
  return hidden == PS_HID ? (int64_t)(2 * num_levels * PS_HID + 2 * PS_HID + 1) : -1;
}

static int check_density_args(int64_t R, int S, int num_levels, int log2_T, int hidden) {
  NSK_REQUIRE(R >= 0 && S >= 1, "nsk_proposal_density: bad shape");
  NSK_REQUIRE(num_levels >= 1 && num_levels <= PS_MAX_LEVELS, "nsk_proposal_density: 1 <= num_levels <= 8");
  NSK_REQUIRE(log2_T >= 1 && log2_T <= 24, "nsk_proposal_density: 1 <= log2_T <= 24");
  NSK_REQUIRE(hidden == PS_HID, "nsk_proposal_density: hidden_dim must be 16");
  return 0;
}

extern "C" int nsk_proposal_density_fwd(const float* origins, const float* dirs, const float* near, const float* far,
                                        const float* bins, int64_t R, int S, const float* table, const float* scalings,
                                        int num_levels, int log2_T, const float* mlp, int hidden, float* density, void* stream) {
  if (check_density_args(R, S, num_levels, log2_T, hidden)) return 1;
  if (R == 0) return 0;
  NSK_REQUIRE(origins && table && scalings && mlp && density, "nsk_proposal_density_fwd: null pointer");
  NSK_REQUIRE(dirs == nullptr || (near && far && bins), "nsk_proposal_density_fwd: ray mode needs near, far, bins");
  if (R == 0) return 0;
  proposal_density_fwd_kernel<<<grid_for(R * S, 256), 256, 0, as_stream(stream)>>>(
      origins, dirs, near, far, bins, R, S, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, mlp, density);
  return check_launch("proposal_density_fwd_kernel");
}

extern "C" int nsk_proposal_density_bwd(const float* origins, const float* dirs, const float* near, const float* far,
                                        const float* bins, int64_t R, int S, const float* table, const float* scalings,
                                        int num_levels, int log2_T, const float* mlp, int hidden, const float* g_density,
                                        float* d_table, float* d_mlp, void* stream) {
  if (check_density_args(R, S, num_levels, log2_T, hidden)) return 1;
  if (R == 0) return 0;
  NSK_REQUIRE(origins && table && scalings && mlp && g_density && d_table && d_mlp, "nsk_proposal_density_bwd: null pointer");
  NSK_REQUIRE(dirs == nullptr || (near && far && bins), "nsk_proposal_density_bwd: ray mode needs near, far, bins");
  if (R == 0) return 0;
  proposal_density_bwd_kernel<<<grid_for(R * S, 256, 148 * 4), 256, 0, as_stream(stream)>>>(
      origins, dirs, near, far, bins, R, S, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, mlp, g_density,
      d_table, d_mlp);
  return check_launch("proposal_density_bwd_kernel");
}

extern "C" int nsk_uniform_bins(const float* base, const float* jitter, int64_t R, int S, float* out, void* stream) {
  NSK_REQUIRE(R >= 0 && S >= 1, "nsk_uniform_bins: bad shape");
  if (R == 0) return 0;
  NSK_REQUIRE(base && out, "nsk_uniform_bins: null pointer");
  const int64_t n = R * (S + 1);
  uniform_bins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(base, jitter, R, S, out);
  return check_launch("uniform_bins_kernel");
}

extern "C" int nsk_pdf_resample(const float* bins, const float* density, const float* weights_in, const float* near,
                                const float* far, int64_t R, int S, int N, float anneal, float histogram_padding, float eps,
                                const float* u_base, float u_half, const float* jitter, float* weights_out, float* new_bins,
                                float* new_euclid, void* stream) {
  NSK_REQUIRE(R >= 0 && S >= 1 && S <= PS_MAX_S && N >= 1, "nsk_pdf_resample: 1 <= S <= 512, N >= 1");
  if (R == 0) return 0;
  NSK_REQUIRE(bins && near && far && u_base && new_bins, "nsk_pdf_resample: null pointer");
  NSK_REQUIRE((density != nullptr) != (weights_in != nullptr), "nsk_pdf_resample: give density OR weights_in");
  const size_t smem = (size_t)PR_WARPS * 2 * (S + 1) * sizeof(float);
  pdf_resample_kernel<<<grid_for(R, PR_WARPS, 148 * 16), PR_WARPS * 32, smem, as_stream(stream)>>>(
      bins, density, weights_in, near, far, R, S, N, anneal, histogram_padding, eps, u_base, u_half, jitter, weights_out, new_bins,
      new_euclid);
  return check_launch("pdf_resample_kernel");
}

extern "C" int nsk_density_weights_bwd(const float* bins, const float* density, const float* near, const float* far, int64_t R,
                                       int S, const float* g_weights, float* g_density, void* stream) {
  NSK_REQUIRE(R >= 0 && S >= 1 && S <= PS_MAX_S, "nsk_density_weights_bwd: 1 <= S <= 512");
  if (R == 0) return 0;
  NSK_REQUIRE(bins && density && near && far && g_weights && g_density, "nsk_density_weights_bwd: null pointer");
  density_weights_bwd_kernel<<<grid_for(R, PR_WARPS, 148 * 16), PR_WARPS * 32, 0, as_stream(stream)>>>(bins, density, near, far, R, S,
                                                                                                     g_weights, g_density);
  return check_launch("density_weights_bwd_kernel");
}

extern "C" int nsk_interlevel_loss(const float* c, const float* w, int Sf, const float* cp, const float* wp, int Sp, int64_t R,
                                   float* loss_ray, float* g_wp, void* stream) {
  NSK_REQUIRE(R >= 0 && Sf >= 1 && Sp >= 1 && Sp <= 4096, "nsk_interlevel_loss: bad shape");
  if (R == 0) return 0;
  NSK_REQUIRE(c && w && cp && wp && loss_ray, "nsk_interlevel_loss: null pointer");
  const size_t smem = (size_t)PR_WARPS * 3 * (Sp + 1) * sizeof(float);
  interlevel_loss_kernel<<<grid_for(R, PR_WARPS, 148 * 16), PR_WARPS * 32, smem, as_stream(stream)>>>(c, w, Sf, cp, wp, Sp, R, loss_ray,
                                                                                                    g_wp);
  return check_launch("interlevel_loss_kernel");
}

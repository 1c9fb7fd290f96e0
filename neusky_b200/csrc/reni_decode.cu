// RENI++ radiance table [K latent codes x D directions] -> HDR RGB.
// Replaces RENIField.get_outputs + unnormalise as driven by NeuSkyFactoModel.sample_illumination
// (ns_reni/reni/illumination_fields/reni_illumination_field.py:198-246, 493-573;
//  ns_reni/reni/field_components/transformer_decoder.py:21-155; vn_layers.py:191-246, 404-419;
//  base_spherical_field.py:143-154; neusky/models/neusky_model.py:445-551).
//
// The decoder's "attention" has a single key/value token, so softmax == 1 and each layer's
// attention output is fc_out(value(cond)) -- a per-latent-code constant.  Kernel 1
// computes those 6 x [H] vectors once per latent code; kernel 2 runs the per-(code, direction)
// residual MLP.  The reference instead expands the latent to [K*D,100,3] and recomputes the
// per-code part for every row.  fp32 throughout (the table is K*D*0.5 MFLOP: negligible work).
#include "reni_common.cuh"

namespace nsk {

// ---- kernel 1: per latent code ----------------------------------------------------------------
__global__ void __launch_bounds__(RENI_H)
reni_prep_kernel(const float* __restrict__ latents, const float* __restrict__ rotation, const float* __restrict__ W,
                 ReniLayout y, float* __restrict__ zxy_out /*[K,L,2] rotated xy*/, float* __restrict__ attn /*[K,NL,H]*/) {
  __shared__ float cond[RENI_MAX_L * 3];
  __shared__ float v[RENI_H];
  const int k = blockIdx.x, t = threadIdx.x;
  const float* Z = latents + (int64_t)k * y.L * 3;
  const float* vn = W + y.vn;
  for (int l = t; l < y.L; l += blockDim.x) {
    float z0 = Z[l * 3], z1 = Z[l * 3 + 1], z2 = Z[l * 3 + 2];
    if (rotation) {  // Z @ R  (reni_illumination_field.py:517-519)
      const float r0 = z0 * rotation[0] + z1 * rotation[3] + z2 * rotation[6];
      const float r1 = z0 * rotation[1] + z1 * rotation[4] + z2 * rotation[7];
      const float r2 = z0 * rotation[2] + z1 * rotation[5] + z2 * rotation[8];
      z0 = r0; z1 = r1; z2 = r2;
    }
    zxy_out[((int64_t)k * y.L + l) * 2] = z0;
    zxy_out[((int64_t)k * y.L + l) * 2 + 1] = z1;
    float c0, c1;
    vn_invariant_xy<float>(z0, z1, vn, c0, c1);
    cond[l * 3 + 0] = c0;
    cond[l * 3 + 1] = c1;
    cond[l * 3 + 2] = z2;  // invariant z component (reni_illumination_field.py:228,244)
  }
  __syncthreads();
  for (int i = 0; i < y.NL; ++i) {
    const float* Wl = W + y.layer0 + (int64_t)i * y.layer_stride;
    float a = Wl[y.o_val_b + t];
    for (int c = 0; c < y.c_in; ++c) a += cond[c] * Wl[y.o_val_wt + (int64_t)c * y.H + t];
    v[t] = a;
    __syncthreads();
    float b = Wl[y.o_out_b + t];
    for (int c = 0; c < y.H; ++c) b += v[c] * Wl[y.o_out_wt + (int64_t)c * y.H + t];
    attn[((int64_t)k * y.NL + i) * y.H + t] = b;
    __syncthreads();
  }
}

// ---- kernel 2: per (latent code, direction) row ---------------------------------------------
__device__ __forceinline__ void block_layernorm(float (&x)[RENI_ROWS], const float* gw, const float* gb, int t,
                                                float (*red)[RENI_H / 32][2]) {
  // LayerNorm over the H = 128 threads of the block, for RENI_ROWS rows at once (eps 1e-5, biased var)
  const int warp = t >> 5, lane = t & 31;
  float mean[RENI_ROWS], rstd[RENI_ROWS];
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) {
    const float s = warp_sum(x[r]);
    if (lane == 0) red[r][warp][0] = s;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) mean[r] = (red[r][0][0] + red[r][1][0] + red[r][2][0] + red[r][3][0]) * (1.0f / RENI_H);
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) {
    const float d = x[r] - mean[r];
    const float s = warp_sum(d * d);
    if (lane == 0) red[r][warp][1] = s;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) {
    const float var = (red[r][0][1] + red[r][1][1] + red[r][2][1] + red[r][3][1]) * (1.0f / RENI_H);
    rstd[r] = rsqrtf(var + 1e-5f);
    x[r] = (x[r] - mean[r]) * rstd[r] * gw[t] + gb[t];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RENI_H)
reni_rows_kernel(const float* __restrict__ dirs, int64_t D, const int* __restrict__ row_cam, const float* __restrict__ zxy,
                 const float* __restrict__ scale, const float* __restrict__ attn, const float* __restrict__ W, ReniLayout y,
                 int log_domain, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* pe = sm;                                  // [d_in][RENI_ROWS]
  float* act = sm + (size_t)y.d_in * RENI_ROWS;    // [H][RENI_ROWS]
  __shared__ float red[RENI_ROWS][RENI_H / 32][2];
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * RENI_ROWS;
  const int Lp2 = y.L + 2;
  const float TWO_PI = 6.283185307179586f, HALF_PI = 1.5707963267948966f;
  // latent code of each row: the table mode (row_cam == NULL) decodes code blockIdx.y in every direction; the per-row mode
  // decodes direction d with code row_cam[d] (background radiance along camera rays, neusky_model.py:535-549)
  int kr[RENI_ROWS];
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) kr[r] = row_cam ? row_cam[min(row0 + r, D - 1)] : (int)blockIdx.y;
  // directional input + NeRF PE (2 freqs {1,4}, include_input appended) [NS-mem A.2]
  for (int e = t; e < Lp2 * RENI_ROWS; e += blockDim.x) {
    const int j = e / RENI_ROWS, r = e % RENI_ROWS;
    const int64_t d = min(row0 + r, D - 1);
    const int k = row_cam ? row_cam[d] : (int)blockIdx.y;
    const float dx = dirs[d * 3], dy = dirs[d * 3 + 1], dz = dirs[d * 3 + 2];
    float xin;
    if (j < y.L) xin = zxy[((int64_t)k * y.L + j) * 2] * dx + zxy[((int64_t)k * y.L + j) * 2 + 1] * dy;
    else if (j == y.L) xin = dz;
    else xin = sqrtf(dx * dx + dy * dy);
    const float s = TWO_PI * xin;
    pe[(j * 2 + 0) * RENI_ROWS + r] = sinf(s * 1.0f);
    pe[(j * 2 + 1) * RENI_ROWS + r] = sinf(s * 4.0f);
    pe[(2 * Lp2 + j * 2 + 0) * RENI_ROWS + r] = sinf(s * 1.0f + HALF_PI);
    pe[(2 * Lp2 + j * 2 + 1) * RENI_ROWS + r] = sinf(s * 4.0f + HALF_PI);
    pe[(4 * Lp2 + j) * RENI_ROWS + r] = xin;
  }
  __syncthreads();
  float x[RENI_ROWS];
  {
    const float b = W[y.res_b + t];
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) x[r] = b;
    const float* wt = W + y.res_wt;
    for (int c = 0; c < y.d_in; ++c) {
      const float w = wt[(int64_t)c * y.H + t];
      const float4 a0 = *reinterpret_cast<const float4*>(pe + c * RENI_ROWS);
      const float4 a1 = *reinterpret_cast<const float4*>(pe + c * RENI_ROWS + 4);
      x[0] += a0.x * w; x[1] += a0.y * w; x[2] += a0.z * w; x[3] += a0.w * w;
      x[4] += a1.x * w; x[5] += a1.y * w; x[6] += a1.z * w; x[7] += a1.w * w;
    }
  }
  for (int i = 0; i < y.NL; ++i) {
    const float* Wl = W + y.layer0 + (int64_t)i * y.layer_stride;
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) x[r] += attn[((int64_t)kr[r] * y.NL + i) * y.H + t];
    block_layernorm(x, Wl + y.o_n1w, Wl + y.o_n1b, t, red);   // out1 = LN(attn + x)
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = x[r];
    __syncthreads();
    float h[RENI_ROWS];
    {
      const float b = Wl[y.o_f0_b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) h[r] = b;
      for (int c = 0; c < y.H; ++c) {
        const float w = Wl[y.o_f0_wt + (int64_t)c * y.H + t];
        const float4 a0 = *reinterpret_cast<const float4*>(act + c * RENI_ROWS);
        const float4 a1 = *reinterpret_cast<const float4*>(act + c * RENI_ROWS + 4);
        h[0] += a0.x * w; h[1] += a0.y * w; h[2] += a0.z * w; h[3] += a0.w * w;
        h[4] += a1.x * w; h[5] += a1.y * w; h[6] += a1.z * w; h[7] += a1.w * w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = fmaxf(h[r], 0.f);
    __syncthreads();
    {
      const float b = Wl[y.o_f2_b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) h[r] = b;
      for (int c = 0; c < y.H; ++c) {
        const float w = Wl[y.o_f2_wt + (int64_t)c * y.H + t];
        const float4 a0 = *reinterpret_cast<const float4*>(act + c * RENI_ROWS);
        const float4 a1 = *reinterpret_cast<const float4*>(act + c * RENI_ROWS + 4);
        h[0] += a0.x * w; h[1] += a0.y * w; h[2] += a0.z * w; h[3] += a0.w * w;
        h[4] += a1.x * w; h[5] += a1.y * w; h[6] += a1.z * w; h[7] += a1.w * w;
      }
    }
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) x[r] += h[r];
    block_layernorm(x, Wl + y.o_n2w, Wl + y.o_n2b, t, red);   // out2 = LN(fc + out1)
  }
  // final 128 -> 3 projection, + scale in the log domain, exp
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = x[r];
  __syncthreads();
  if (t < 3 * RENI_ROWS) {
    const int r = t / 3, c = t % 3;
    if (row0 + r < D) {
      float o = W[y.fc_b + c];
      for (int j = 0; j < y.H; ++j) o += act[j * RENI_ROWS + r] * W[y.fc_w + (int64_t)c * y.H + j];
      const int k = row_cam ? row_cam[row0 + r] : (int)blockIdx.y;
      if (scale) {
        const float s = expf(scale[k]);
        o = log_domain ? (o + logf(s)) : (o * s);
      }
      // log_domain: 0 = linear model; 1 = log-domain model, unnormalised radiance exp(o) (BaseRENIField.unnormalise folded in);
      // 2 = log-domain model, the raw log value as RENIField.forward returns it (no exp -> log round trip, which underflows to -inf)
      out[((row_cam ? 0 : (int64_t)k * D) + row0 + r) * 3 + c] = log_domain == 1 ? expf(o) : o;
    }
  }
}

}  // namespace nsk

extern "C" int64_t nsk_reni_weights_floats(int latent_dim, int hidden, int num_layers) {
  return nsk::reni_layout(latent_dim, hidden, num_layers).total;
}

int nsk::reni_launch_prep(const float* latents, const float* rotation, const float* W, ReniLayout y, int64_t K, float* zxy, float* attn, cudaStream_t st) {
  nsk::reni_prep_kernel<<<(unsigned)K, nsk::RENI_H, 0, st>>>(latents, rotation, W, y, zxy, attn);
  return nsk::check_launch("reni_prep_kernel");
}

static int reni_decode_launch(const char* what, const float* dirs, int64_t D, const int* row_cam, const float* latents, const float* scale,
                              int64_t K, const float* rotation, const float* weights, int latent_dim, int hidden, int num_layers,
                              int log_domain, float* workspace, float* out, void* stream) {
  NSK_REQUIRE(hidden == nsk::RENI_H, "nsk_reni_decode: hidden_features must be 128");
  NSK_REQUIRE(latent_dim >= 1 && latent_dim <= nsk::RENI_MAX_L, "nsk_reni_decode: latent_dim out of range");
  if (K == 0 || D == 0) return 0;
  NSK_REQUIRE(dirs && latents && weights && workspace && out, "nsk_reni_decode: null pointer");
  NSK_REQUIRE(K <= 65535, "nsk_reni_decode: too many latent codes for one launch");
  const nsk::ReniLayout y = nsk::reni_layout(latent_dim, hidden, num_layers);
  float* attn = workspace;                                  // [K, NL, H]
  float* zxy = workspace + K * num_layers * (int64_t)hidden;  // [K, L, 2]
  cudaStream_t st = nsk::as_stream(stream);
  if (int e = nsk::reni_launch_prep(latents, rotation, weights, y, K, zxy, attn, st)) return e;
  const size_t smem = ((size_t)y.d_in + hidden) * nsk::RENI_ROWS * sizeof(float);
  dim3 grid((unsigned)((D + nsk::RENI_ROWS - 1) / nsk::RENI_ROWS), row_cam ? 1u : (unsigned)K);
  nsk::reni_rows_kernel<<<grid, nsk::RENI_H, smem, st>>>(dirs, D, row_cam, zxy, scale, attn, weights, y, log_domain, out);
  return nsk::check_launch(what);
}

extern "C" int nsk_reni_decode_fwd(const float* dirs, int64_t D, const float* latents, const float* scale, int64_t K,
                                   const float* rotation, const float* weights, int latent_dim, int hidden, int num_layers,
                                   int log_domain, float* workspace, float* out, void* stream) {
  return reni_decode_launch("nsk_reni_decode_fwd", dirs, D, nullptr, latents, scale, K, rotation, weights, latent_dim, hidden, num_layers,
                            log_domain, workspace, out, stream);
}

extern "C" int nsk_reni_decode_rows_fwd(const float* dirs, const int* row_cam, int64_t N, const float* latents, const float* scale, int64_t K,
                                        const float* rotation, const float* weights, int latent_dim, int hidden, int num_layers,
                                        int log_domain, float* workspace, float* out, void* stream) {
  NSK_REQUIRE(row_cam != nullptr || N == 0, "nsk_reni_decode_rows_fwd: null row_cam");
  return reni_decode_launch("nsk_reni_decode_rows_fwd", dirs, N, row_cam, latents, scale, K, rotation, weights, latent_dim, hidden, num_layers,
                            log_domain, workspace, out, stream);
}

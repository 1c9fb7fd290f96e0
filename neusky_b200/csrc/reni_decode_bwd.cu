// RENI++ radiance backward: d loss / d latent codes and d loss / d scale, decoder weights frozen.
//
// The reference trains one latent code Z [L,3] and one scale per image through a FIXED RENI++ decoder
// (fixed_decoder=True, neusky/configs/neusky_config.py:94; RENIField.hold_decoder_fixed,
// ns_reni/reni/illumination_fields/reni_illumination_field.py:157-196): torch autograd walks back through
// unnormalise (exp), the + scale, the transformer decoder, the NeRF encoding of the SO(2)-invariant direction
// input and the VN-invariant conditioning (reni_illumination_field.py:198-246, 493-573;
// ns_reni/reni/field_components/transformer_decoder.py:21-155; vn_layers.py:191-246, 404-419).
//
// Here the forward is recomputed per block of 8 rows with its LayerNorm / ReLU state kept in shared memory
// (nothing was saved by the forward), the row gradient is pushed back through the transposed products, and the
// two places where the latent code enters are reduced with atomics:
//     d attn[k, layer, :]  (the per-code attention constants)     -> reni_prep_bwd_kernel -> d cond -> d Z
//     d zxy[k, l, :]       (the Z_xy . d_xy inner products)                                        -> d Z
// fp32 throughout.  Work: 2 x 262,272 MAC per row (recompute + transposed pass); the per-code part is K blocks.
#include "reni_common.cuh"

namespace nsk {

// y[t] = b[t] + sum_c in[c][r] * wt[c*H + t]  for RENI_ROWS rows (in: shared [C][RENI_ROWS])
__device__ __forceinline__ void rows_matvec(float (&acc)[RENI_ROWS], const float* __restrict__ wt, const float* in, int C, int t) {
  for (int c = 0; c < C; ++c) {
    const float w = wt[(int64_t)c * RENI_H + t];
    const float4 a0 = *reinterpret_cast<const float4*>(in + c * RENI_ROWS);
    const float4 a1 = *reinterpret_cast<const float4*>(in + c * RENI_ROWS + 4);
    acc[0] += a0.x * w; acc[1] += a0.y * w; acc[2] += a0.z * w; acc[3] += a0.w * w;
    acc[4] += a1.x * w; acc[5] += a1.y * w; acc[6] += a1.z * w; acc[7] += a1.w * w;
  }
}
// transposed product with the [out][in] copy: din[c][r] = sum_t g[t][r] * w[t*ld + c]   (c = this thread's column)
__device__ __forceinline__ void rows_matvec_t(float (&acc)[RENI_ROWS], const float* __restrict__ w, int ld, const float* g, int c) {
  for (int t = 0; t < RENI_H; ++t) {
    const float wv = w[(int64_t)t * ld + c];
    const float4 a0 = *reinterpret_cast<const float4*>(g + t * RENI_ROWS);
    const float4 a1 = *reinterpret_cast<const float4*>(g + t * RENI_ROWS + 4);
    acc[0] += a0.x * wv; acc[1] += a0.y * wv; acc[2] += a0.z * wv; acc[3] += a0.w * wv;
    acc[4] += a1.x * wv; acc[5] += a1.y * wv; acc[6] += a1.z * wv; acc[7] += a1.w * wv;
  }
}

// shared-memory plan (floats): pe [d_in][8] | act [H][8] | per layer { xh1 [H][8], hr [H][8], xh2 [H][8], rstd1 [8], rstd2 [8] }
__host__ __device__ inline size_t reni_bwd_smem_floats(int d_in, int NL) {
  return (size_t)d_in * RENI_ROWS + (size_t)RENI_H * RENI_ROWS + (size_t)NL * (3 * RENI_H * RENI_ROWS + 2 * RENI_ROWS);
}

__global__ void __launch_bounds__(RENI_H)
reni_rows_bwd_kernel(const float* __restrict__ dirs, int64_t D, const int* __restrict__ row_cam, const float* __restrict__ zxy,
                     const float* __restrict__ attn, const float* __restrict__ W, ReniLayout y, const float* __restrict__ WB, ReniBwdLayout yb,
                     int log_domain, int has_scale, const float* __restrict__ scale, const float* __restrict__ out, const float* __restrict__ g_out,
                     float* __restrict__ d_attn /*[K,NL,H]*/, float* __restrict__ d_zxy /*[K,L,2]*/, float* __restrict__ d_scale /*[K]*/) {
  extern __shared__ float sm[];
  float* pe = sm;
  float* act = pe + (size_t)y.d_in * RENI_ROWS;
  float* lay = act + RENI_H * RENI_ROWS;
  const int LAY = 3 * RENI_H * RENI_ROWS + 2 * RENI_ROWS;
  __shared__ float red[RENI_ROWS][RENI_H / 32][2];
  __shared__ float dout[RENI_ROWS][3];
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * RENI_ROWS;
  const int Lp2 = y.L + 2;
  const float TWO_PI = 6.283185307179586f, HALF_PI = 1.5707963267948966f;
  int kr[RENI_ROWS];
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) kr[r] = row_cam ? row_cam[min(row0 + r, D - 1)] : (int)blockIdx.y;

  // ---------------- forward recompute (same op order as reni_rows_kernel) ----------------
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
    rows_matvec(x, W + y.res_wt, pe, y.d_in, t);
  }
  for (int i = 0; i < y.NL; ++i) {
    const float* Wl = W + y.layer0 + (int64_t)i * y.layer_stride;
    float* L_ = lay + (size_t)i * LAY;
    float* xh1 = L_; float* hr = L_ + RENI_H * RENI_ROWS; float* xh2 = L_ + 2 * RENI_H * RENI_ROWS;
    float* rs1 = L_ + 3 * RENI_H * RENI_ROWS; float* rs2 = rs1 + RENI_ROWS;
    float sa[RENI_ROWS], sb[RENI_ROWS], dv[RENI_ROWS];
    // out1 = LN1(attn + x)
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) x[r] += attn[((int64_t)kr[r] * y.NL + i) * y.H + t];
    block_rowsum2(x, x, sa, sb, t, red);
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dv[r] = x[r] - sa[r] * (1.0f / RENI_H);
    float sq[RENI_ROWS];
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) sq[r] = dv[r] * dv[r];
    block_rowsum2(sq, sq, sa, sb, t, red);
    {
      const float gw = Wl[y.o_n1w + t], gb = Wl[y.o_n1b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) {
        const float rstd = rsqrtf(sa[r] * (1.0f / RENI_H) + 1e-5f);
        const float xh = dv[r] * rstd;
        xh1[t * RENI_ROWS + r] = xh;
        if (t == 0) rs1[r] = rstd;
        x[r] = xh * gw + gb;
        act[t * RENI_ROWS + r] = x[r];
      }
    }
    __syncthreads();
    float h[RENI_ROWS];
    {
      const float b = Wl[y.o_f0_b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) h[r] = b;
      rows_matvec(h, Wl + y.o_f0_wt, act, RENI_H, t);
    }
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) hr[t * RENI_ROWS + r] = fmaxf(h[r], 0.f);
    __syncthreads();
    {
      const float b = Wl[y.o_f2_b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) h[r] = b;
      rows_matvec(h, Wl + y.o_f2_wt, hr, RENI_H, t);
    }
    // out2 = LN2(fc + out1)
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) x[r] += h[r];
    block_rowsum2(x, x, sa, sb, t, red);
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dv[r] = x[r] - sa[r] * (1.0f / RENI_H);
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) sq[r] = dv[r] * dv[r];
    block_rowsum2(sq, sq, sa, sb, t, red);
    {
      const float gw = Wl[y.o_n2w + t], gb = Wl[y.o_n2b + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) {
        const float rstd = rsqrtf(sa[r] * (1.0f / RENI_H) + 1e-5f);
        const float xh = dv[r] * rstd;
        xh2[t * RENI_ROWS + r] = xh;
        if (t == 0) rs2[r] = rstd;
        x[r] = xh * gw + gb;
      }
    }
  }

  // ---------------- backward ----------------
  // out = exp(o + scale) (log domain) or o * exp(scale): d o = g_out * d out / d o, taken from the forward's own output
  if (t < 3 * RENI_ROWS) {
    const int r = t / 3, c = t % 3;
    float go = 0.f;
    if (row0 + r < D) {
      const int64_t oi = ((row_cam ? 0 : (int64_t)kr[0] * D) + row0 + r) * 3 + c;   // table mode: every row of the block has code blockIdx.y
      const int k = row_cam ? row_cam[row0 + r] : (int)blockIdx.y;
      const float g = g_out[oi], ov = out[oi];
      if (log_domain) {
        go = g * ov;                       // d exp(o)/d o = exp(o); d o / d scale = 1 (o + log(exp(scale)))
        if (has_scale) atomicAdd(d_scale + k, go);
      } else {
        const float s = has_scale ? expf(scale[k]) : 1.0f;
        go = g * s;                        // out = o * s
        if (has_scale) atomicAdd(d_scale + k, g * ov);   // d (o s)/d scale = o s
      }
    }
    dout[r][c] = go;
  }
  __syncthreads();
  float dx[RENI_ROWS];
  {
    const float w0 = W[y.fc_w + t], w1 = W[y.fc_w + y.H + t], w2 = W[y.fc_w + 2 * y.H + t];
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dx[r] = dout[r][0] * w0 + dout[r][1] * w1 + dout[r][2] * w2;
  }
  for (int i = y.NL - 1; i >= 0; --i) {
    const float* Wl = W + y.layer0 + (int64_t)i * y.layer_stride;
    const float* WBl = WB + yb.layer0 + (int64_t)i * yb.layer_stride;
    float* L_ = lay + (size_t)i * LAY;
    const float* xh1 = L_; const float* hr = L_ + RENI_H * RENI_ROWS; const float* xh2 = L_ + 2 * RENI_H * RENI_ROWS;
    const float* rs1 = L_ + 3 * RENI_H * RENI_ROWS; const float* rs2 = rs1 + RENI_ROWS;
    float a[RENI_ROWS], b[RENI_ROWS], sa[RENI_ROWS], sb[RENI_ROWS];
    // LN2 backward: dx holds d out2
    {
      const float gw = Wl[y.o_n2w + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) { a[r] = dx[r] * gw; b[r] = a[r] * xh2[t * RENI_ROWS + r]; }
    }
    block_rowsum2(a, b, sa, sb, t, red);
    float du[RENI_ROWS];   // d (fc + out1)
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r)
      du[r] = rs2[r] * (a[r] - sa[r] * (1.0f / RENI_H) - xh2[t * RENI_ROWS + r] * sb[r] * (1.0f / RENI_H));
    // fc.2 transposed, ReLU mask
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = du[r];
    __syncthreads();
    float dh[RENI_ROWS];
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dh[r] = 0.f;
    rows_matvec_t(dh, WBl + yb.o_f2_w, RENI_H, act, t);
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dh[r] = hr[t * RENI_ROWS + r] > 0.f ? dh[r] : 0.f;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = dh[r];
    __syncthreads();
    // fc.0 transposed, + residual: d out1
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) dx[r] = du[r];
    rows_matvec_t(dx, WBl + yb.o_f0_w, RENI_H, act, t);
    // LN1 backward
    {
      const float gw = Wl[y.o_n1w + t];
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) { a[r] = dx[r] * gw; b[r] = a[r] * xh1[t * RENI_ROWS + r]; }
    }
    block_rowsum2(a, b, sa, sb, t, red);
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r)
      dx[r] = rs1[r] * (a[r] - sa[r] * (1.0f / RENI_H) - xh1[t * RENI_ROWS + r] * sb[r] * (1.0f / RENI_H));
    // d (attn + x): the attention constant of this row's code gets the same gradient as x
    if (row_cam) {
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r)
        if (row0 + r < D) atomicAdd(d_attn + ((int64_t)kr[r] * y.NL + i) * y.H + t, dx[r]);
    } else {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r)
        if (row0 + r < D) s += dx[r];
      atomicAdd(d_attn + ((int64_t)kr[0] * y.NL + i) * y.H + t, s);
    }
  }
  // residual projection transposed -> d pe -> d xin -> d zxy
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) act[t * RENI_ROWS + r] = (row0 + r < D) ? dx[r] : 0.f;
  __syncthreads();
  // thread handles input features j = t (only j < L feed the latent): 5 encoded columns each
  if (t < y.L) {
    const int j = t;
    const int cols[5] = {j * 2 + 0, j * 2 + 1, 2 * Lp2 + j * 2 + 0, 2 * Lp2 + j * 2 + 1, 4 * Lp2 + j};
    float dp[5][RENI_ROWS];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
#pragma unroll
      for (int r = 0; r < RENI_ROWS; ++r) dp[q][r] = 0.f;
      rows_matvec_t(dp[q], WB + yb.res_w, y.d_in, act, cols[q]);
    }
    float gx = 0.f, gy = 0.f;   // table mode: all rows share the code -> one atomic pair per thread
#pragma unroll
    for (int r = 0; r < RENI_ROWS; ++r) {
      if (row0 + r >= D) continue;
      // d sin(2 pi f x)/dx = 2 pi f cos(.) ; d cos/dx = -2 pi f sin(.) ; the cos / sin values are the encoded columns themselves
      const float s1 = pe[cols[0] * RENI_ROWS + r], s4 = pe[cols[1] * RENI_ROWS + r];
      const float c1 = pe[cols[2] * RENI_ROWS + r], c4 = pe[cols[3] * RENI_ROWS + r];
      const float dxin = TWO_PI * (dp[0][r] * c1 - dp[2][r] * s1) + 4.0f * TWO_PI * (dp[1][r] * c4 - dp[3][r] * s4) + dp[4][r];
      const int64_t d = row0 + r;
      const float ddx = dirs[d * 3], ddy = dirs[d * 3 + 1];
      if (row_cam) {
        atomicAdd(d_zxy + ((int64_t)kr[r] * y.L + j) * 2 + 0, dxin * ddx);
        atomicAdd(d_zxy + ((int64_t)kr[r] * y.L + j) * 2 + 1, dxin * ddy);
      } else {
        gx += dxin * ddx; gy += dxin * ddy;
      }
    }
    if (!row_cam) {
      atomicAdd(d_zxy + ((int64_t)kr[0] * y.L + j) * 2 + 0, gx);
      atomicAdd(d_zxy + ((int64_t)kr[0] * y.L + j) * 2 + 1, gy);
    }
  }
}

// ---- per latent code: d attn -> d cond -> d Z; d zxy -> d Z; un-rotate ----------------------------------------
__global__ void __launch_bounds__(RENI_H)
reni_prep_bwd_kernel(const float* __restrict__ latents, const float* __restrict__ rotation, const float* __restrict__ W, ReniLayout y,
                     const float* __restrict__ WB, ReniBwdLayout yb, const float* __restrict__ d_attn, const float* __restrict__ d_zxy,
                     float* __restrict__ d_latents /*[K,L,3], accumulated*/) {
  __shared__ float dcond[RENI_MAX_L * 3];
  __shared__ float g[RENI_H], dv[RENI_H];
  const int k = blockIdx.x, t = threadIdx.x;
  for (int c = t; c < y.c_in; c += blockDim.x) dcond[c] = 0.f;
  __syncthreads();
  for (int i = 0; i < y.NL; ++i) {
    const float* WBl = WB + yb.layer0 + (int64_t)i * yb.layer_stride;
    g[t] = d_attn[((int64_t)k * y.NL + i) * y.H + t];
    __syncthreads();
    // attn = out_b + v . out_wt ; v = val_b + cond . val_wt
    float s = 0.f;
    for (int o = 0; o < y.H; ++o) s += g[o] * WBl[yb.o_out_w + (int64_t)o * y.H + t];
    dv[t] = s;
    __syncthreads();
    for (int c = t; c < y.c_in; c += blockDim.x) {
      float a = 0.f;
      for (int o = 0; o < y.H; ++o) a += dv[o] * WBl[yb.o_val_w + (int64_t)o * y.c_in + c];
      dcond[c] += a;
    }
    __syncthreads();
  }
  const float* Z = latents + (int64_t)k * y.L * 3;
  const float* vn = W + y.vn;
  for (int l = t; l < y.L; l += blockDim.x) {
    float z0 = Z[l * 3], z1 = Z[l * 3 + 1], z2 = Z[l * 3 + 2];
    if (rotation) {
      const float r0 = z0 * rotation[0] + z1 * rotation[3] + z2 * rotation[6];
      const float r1 = z0 * rotation[1] + z1 * rotation[4] + z2 * rotation[7];
      z0 = r0; z1 = r1;
    }
    Dual2 c0, c1;
    vn_invariant_xy<Dual2>(Dual2{z0, 1.f, 0.f}, Dual2{z1, 0.f, 1.f}, vn, c0, c1);
    const float g0 = dcond[l * 3 + 0], g1 = dcond[l * 3 + 1];
    // gradient w.r.t. the rotated latent row
    const float dr0 = g0 * c0.a + g1 * c1.a + d_zxy[((int64_t)k * y.L + l) * 2 + 0];
    const float dr1 = g0 * c0.b + g1 * c1.b + d_zxy[((int64_t)k * y.L + l) * 2 + 1];
    const float dr2 = dcond[l * 3 + 2];
    float o0 = dr0, o1 = dr1, o2 = dr2;
    if (rotation) {   // Zr = Z @ R  ->  dZ = dZr @ R^T
      o0 = dr0 * rotation[0] + dr1 * rotation[1] + dr2 * rotation[2];
      o1 = dr0 * rotation[3] + dr1 * rotation[4] + dr2 * rotation[5];
      o2 = dr0 * rotation[6] + dr1 * rotation[7] + dr2 * rotation[8];
    }
    float* dz = d_latents + ((int64_t)k * y.L + l) * 3;
    dz[0] += o0; dz[1] += o1; dz[2] += o2;
  }
}

}  // namespace nsk

extern "C" int64_t nsk_reni_bwd_weights_floats(int latent_dim, int hidden, int num_layers) {
  return nsk::reni_bwd_layout(latent_dim, hidden, num_layers).total;
}

extern "C" int64_t nsk_reni_bwd_workspace_floats(int64_t K, int latent_dim, int hidden, int num_layers) {
  return 2 * (K * num_layers * (int64_t)hidden + K * latent_dim * 2);
}

extern "C" int nsk_reni_decode_bwd(const float* dirs, const int* row_cam, int64_t D, const float* latents, const float* scale, int64_t K,
                                   const float* rotation, const float* weights, const float* weights_bwd, int latent_dim, int hidden,
                                   int num_layers, int log_domain, const float* out, const float* g_out, float* workspace,
                                   float* d_latents, float* d_scale, void* stream) {
  NSK_REQUIRE(hidden == nsk::RENI_H, "nsk_reni_decode_bwd: hidden_features must be 128");
  NSK_REQUIRE(latent_dim >= 1 && latent_dim <= nsk::RENI_MAX_L, "nsk_reni_decode_bwd: latent_dim out of range");
  if (K == 0 || D == 0) return 0;
  NSK_REQUIRE(dirs && latents && weights && weights_bwd && workspace && out && g_out && d_latents, "nsk_reni_decode_bwd: null pointer");
  NSK_REQUIRE((scale == nullptr) == (d_scale == nullptr), "nsk_reni_decode_bwd: scale and d_scale must be given together");
  NSK_REQUIRE(K <= 65535, "nsk_reni_decode_bwd: too many latent codes for one launch");
  const nsk::ReniLayout y = nsk::reni_layout(latent_dim, hidden, num_layers);
  const nsk::ReniBwdLayout yb = nsk::reni_bwd_layout(latent_dim, hidden, num_layers);
  const int64_t n_attn = K * num_layers * (int64_t)hidden, n_zxy = K * latent_dim * 2;
  float* attn = workspace;
  float* zxy = attn + n_attn;
  float* d_attn = zxy + n_zxy;
  float* d_zxy = d_attn + n_attn;
  cudaStream_t st = nsk::as_stream(stream);
  if (cudaMemsetAsync(d_attn, 0, (size_t)(n_attn + n_zxy) * sizeof(float), st) != cudaSuccess) return nsk::fail("nsk_reni_decode_bwd", "memset failed");
  if (int e = nsk::reni_launch_prep(latents, rotation, weights, y, K, zxy, attn, st)) return e;
  const size_t smem = nsk::reni_bwd_smem_floats(y.d_in, num_layers) * sizeof(float);
  static nsk::DeviceOnce once;
  if (int err = nsk::device_once(once, "nsk_reni_decode_bwd: cannot raise the dynamic shared memory limit", nullptr, [] {
        return cudaFuncSetAttribute(nsk::reni_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
      }))
    return err;
  NSK_REQUIRE(smem <= 200 * 1024, "nsk_reni_decode_bwd: too many layers for the shared-memory plan");
  dim3 grid((unsigned)((D + nsk::RENI_ROWS - 1) / nsk::RENI_ROWS), row_cam ? 1u : (unsigned)K);
  nsk::reni_rows_bwd_kernel<<<grid, nsk::RENI_H, smem, st>>>(dirs, D, row_cam, zxy, attn, weights, y, weights_bwd, yb, log_domain, scale != nullptr,
                                                              scale, out, g_out, d_attn, d_zxy, d_scale);
  if (int e = nsk::check_launch("reni_rows_bwd_kernel")) return e;
  nsk::reni_prep_bwd_kernel<<<(unsigned)K, nsk::RENI_H, 0, st>>>(latents, rotation, weights, y, weights_bwd, yb, d_attn, d_zxy, d_latents);
  return nsk::check_launch("reni_prep_bwd_kernel");
}

// RENI++ decoder: weight-blob layouts and the per-latent-row VN-invariant map shared by the forward
// (reni_decode.cu) and backward (reni_decode_bwd.cu) kernels.
#pragma once
#include "nsk_common.cuh"

namespace nsk {

constexpr int RENI_H = 128;
constexpr int RENI_MAX_L = 128;
constexpr int RENI_ROWS = 8;

// forward blob (python: neusky_b200.packing.pack_reni): linear weights transposed to [in][out]
struct ReniLayout {
  int L, H, NL, d_in, c_in;
  int64_t vn, res_wt, res_b, layer0, layer_stride, fc_w, fc_b, total;
  // per-layer offsets relative to layer start
  int64_t o_val_wt, o_val_b, o_out_wt, o_out_b, o_n1w, o_n1b, o_f0_wt, o_f0_b, o_f2_wt, o_f2_b, o_n2w, o_n2b;
};

__host__ __device__ inline ReniLayout reni_layout(int L, int H, int NL) {
  ReniLayout y;
  y.L = L; y.H = H; y.NL = NL;
  y.d_in = (L + 2) * 5;
  y.c_in = L * 3;
  int64_t o = 0;
  y.vn = o; o += 16;  // [proj(1), lin(2), W(4), U(4), pad]
  y.res_wt = o; o += (int64_t)y.d_in * H;
  y.res_b = o; o += H;
  y.layer0 = o;
  int64_t q = 0;
  y.o_val_wt = q; q += (int64_t)y.c_in * H;
  y.o_val_b = q; q += H;
  y.o_out_wt = q; q += (int64_t)H * H;
  y.o_out_b = q; q += H;
  y.o_n1w = q; q += H;
  y.o_n1b = q; q += H;
  y.o_f0_wt = q; q += (int64_t)H * H;
  y.o_f0_b = q; q += H;
  y.o_f2_wt = q; q += (int64_t)H * H;
  y.o_f2_b = q; q += H;
  y.o_n2w = q; q += H;
  y.o_n2b = q; q += H;
  y.layer_stride = q;
  o += q * NL;
  y.fc_w = o; o += 3 * (int64_t)H;
  y.fc_b = o; o += 4;
  y.total = o;
  return y;
}

// backward blob (python: neusky_b200.packing.pack_reni_bwd): the same linear weights in torch's own [out][in]
// layout, which is the coalesced one for the transposed products of the backward pass
struct ReniBwdLayout {
  int64_t res_w /*[H][d_in]*/, layer0, layer_stride, total;
  int64_t o_val_w /*[H][c_in]*/, o_out_w /*[H][H]*/, o_f0_w /*[H][H]*/, o_f2_w /*[H][H]*/;
};

__host__ __device__ inline ReniBwdLayout reni_bwd_layout(int L, int H, int NL) {
  ReniBwdLayout b;
  int64_t o = 0;
  b.res_w = o; o += (int64_t)H * (L + 2) * 5;
  b.layer0 = o;
  int64_t q = 0;
  b.o_val_w = q; q += (int64_t)H * L * 3;
  b.o_out_w = q; q += (int64_t)H * H;
  b.o_f0_w = q; q += (int64_t)H * H;
  b.o_f2_w = q; q += (int64_t)H * H;
  b.layer_stride = q;
  o += q * NL;
  b.total = o;
  return b;
}

// ---- forward-mode pair: value + d/dz0 + d/dz1 (the VN map is R^2 -> R^2 per latent row; its 2x2 Jacobian is all
// the backward pass needs, and running the SAME code on this type keeps forward and backward consistent) ----------
struct Dual2 {
  float v, a, b;
};
__device__ __forceinline__ Dual2 operator+(Dual2 x, Dual2 y) { return {x.v + y.v, x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ Dual2 operator-(Dual2 x, Dual2 y) { return {x.v - y.v, x.a - y.a, x.b - y.b}; }
__device__ __forceinline__ Dual2 operator*(Dual2 x, Dual2 y) { return {x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b}; }
__device__ __forceinline__ Dual2 operator*(float s, Dual2 y) { return {s * y.v, s * y.a, s * y.b}; }
__device__ __forceinline__ Dual2 operator/(Dual2 x, Dual2 y) {
  const float q = x.v / y.v;
  return {q, (x.a - q * y.a) / y.v, (x.b - q * y.b) / y.v};
}
__device__ __forceinline__ float val(float x) { return x; }
__device__ __forceinline__ float val(Dual2 x) { return x.v; }
// sqrt(max(x, eps)): the clamp has zero slope on the clamped side (torch.clamp semantics)
__device__ __forceinline__ float sqrt_clamped(float x, float eps) { return sqrtf(fmaxf(x, eps)); }
__device__ __forceinline__ Dual2 sqrt_clamped(Dual2 x, float eps) {
  if (x.v < eps) return {sqrtf(eps), 0.f, 0.f};
  const float s = sqrtf(x.v);
  return {s, 0.5f * x.a / s, 0.5f * x.b / s};
}

// vn_proj_in (VNLinear(1,1)) -> VNInvariant(dim=1, dim_coor=2): VNLinear(1,2), VNReLU(2), contraction with the input
// (ns_reni/reni/field_components/vn_layers.py:191-246, 404-419; reni_illumination_field.py:219-246).
// vn = [proj(1), lin(2), W(4), U(4)].  (z0, z1) = xy of one (rotated) latent row -> its two invariants.
template <typename T>
__device__ __forceinline__ void vn_invariant_xy(T z0, T z1, const float* __restrict__ vn, T& c0, T& c1) {
  const T x0 = vn[0] * z0, x1 = vn[0] * z1;                                   // VNLinear(1,1): x[c] = w * z[c]
  const T yv[2][2] = {{vn[1] * x0, vn[1] * x1}, {vn[2] * x0, vn[2] * x1}};    // VNLinear(1,2): yv[o][c] = w0[o] * x[c]
  T outv[2][2];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    // VNReLU(2): q = W y, k = U y over the feature index, per coordinate c
    const T q0 = vn[3 + o * 2 + 0] * yv[0][0] + vn[3 + o * 2 + 1] * yv[1][0];
    const T q1 = vn[3 + o * 2 + 0] * yv[0][1] + vn[3 + o * 2 + 1] * yv[1][1];
    const T k0 = vn[7 + o * 2 + 0] * yv[0][0] + vn[7 + o * 2 + 1] * yv[1][0];
    const T k1 = vn[7 + o * 2 + 0] * yv[0][1] + vn[7 + o * 2 + 1] * yv[1][1];
    const T qk = q0 * k0 + q1 * k1;
    const T kn = sqrt_clamped(k0 * k0 + k1 * k1, 1e-6f);
    const T proj = q0 * (k0 / kn) + q1 * (k1 / kn);
    const bool keep = val(qk) >= 0.f;
    outv[o][0] = keep ? q0 : (q0 - proj * k0);
    outv[o][1] = keep ? q1 : (q1 - proj * k1);
  }
  // rearrange '... d e -> ... e d', einsum('b n d i, b n i o -> b n o') with d == 1: inv[o] = sum_i x[i] outv[o][i]
  c0 = x0 * outv[0][0] + x1 * outv[0][1];
  c1 = x0 * outv[1][0] + x1 * outv[1][1];
}

// per-code prologue (reni_decode.cu): rotated Z_xy [K,L,2] and the six attention constants [K,NL,H]
int reni_launch_prep(const float* latents, const float* rotation, const float* W, ReniLayout y, int64_t K, float* zxy, float* attn, cudaStream_t st);

// LayerNorm statistics over the H = 128 threads of a block for RENI_ROWS rows at once
__device__ __forceinline__ void block_rowsum2(const float (&a)[RENI_ROWS], const float (&b)[RENI_ROWS], float (&sa)[RENI_ROWS], float (&sb)[RENI_ROWS],
                                              int t, float (*red)[RENI_H / 32][2]) {
  const int warp = t >> 5, lane = t & 31;
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) {
    const float s0 = warp_sum(a[r]), s1 = warp_sum(b[r]);
    if (lane == 0) { red[r][warp][0] = s0; red[r][warp][1] = s1; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RENI_ROWS; ++r) {
    sa[r] = red[r][0][0] + red[r][1][0] + red[r][2][0] + red[r][3][0];
    sb[r] = red[r][0][1] + red[r][1][1] + red[r][2][1] + red[r][3][1];
  }
  __syncthreads();
}

}  // namespace nsk

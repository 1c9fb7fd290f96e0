// RENI++ decode of MANY rows (the per-ray background radiance of a frame, neusky/models/neusky_model.py:535-549, and large
// direction sets) as a chain of 3xTF32 tensor-core contractions: the row-wise residual MLP of the decoder
// (ns_reni/reni/field_components/transformer_decoder.py:21-155) is 1 + 12 dense layers over [N, 512 | 128] activations, which
// the SIMT kernel of reni_decode.cu runs at ~18 TFLOP/s (one hidden unit per thread, 8 rows per block).  Here the layers run on
// nsk_gemm_tf32_nt (fp32-accurate 3xTF32) and this file supplies the pieces between them:
//   nsk_reni_prep       per latent code: rotated latent xy [K,L,2] and the 6 attention vectors [K,NL,H] (same kernel as the
//                       table path: the single-token attention is a per-code constant)
//   nsk_reni_pe_rows    decoder input rows [N,512]: VN-invariant inner products, d_z, |d_xy| with their NeRF encoding
//                       (reni_illumination_field.py:219-246, 345-348), zero-padded from 510 to the MMA K step
//   nsk_reni_ln_rows    x <- LayerNorm(x + add[code(row)]) * w + b in place, one warp per row (eps 1e-5, biased variance), optionally
//                       followed by a second (add, LayerNorm) in the same pass
// All fp32; the elementwise kernels are one pass over their operand (HBM-bound).
#include "reni_common.cuh"

namespace nsk {

int reni_launch_prep(const float* latents, const float* rotation, const float* W, ReniLayout y, int64_t K, float* zxy, float* attn, cudaStream_t st);

constexpr int RENI_PE_LD = 512;

__global__ void __launch_bounds__(256)
reni_pe_rows_kernel(const float* __restrict__ dirs, const int* __restrict__ row_cam, int64_t N, const float* __restrict__ zxy, int L,
                    float* __restrict__ pe) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int k = row_cam ? row_cam[row] : 0;
  const float dx = dirs[row * 3], dy = dirs[row * 3 + 1], dz = dirs[row * 3 + 2];
  const float TWO_PI = 6.283185307179586f, HALF_PI = 1.5707963267948966f;
  const int Lp2 = L + 2;
  float* pr = pe + row * RENI_PE_LD;
  for (int j = lane; j < Lp2; j += 32) {
    float xin;
    if (j < L) xin = zxy[((int64_t)k * L + j) * 2] * dx + zxy[((int64_t)k * L + j) * 2 + 1] * dy;
    else if (j == L) xin = dz;
    else xin = sqrtf(dx * dx + dy * dy);
    const float s = TWO_PI * xin;
    pr[j * 2 + 0] = sinf(s * 1.0f);
    pr[j * 2 + 1] = sinf(s * 4.0f);
    pr[2 * Lp2 + j * 2 + 0] = sinf(s * 1.0f + HALF_PI);
    pr[2 * Lp2 + j * 2 + 1] = sinf(s * 4.0f + HALF_PI);
    pr[4 * Lp2 + j] = xin;
  }
  for (int c = 5 * Lp2 + lane; c < RENI_PE_LD; c += 32) pr[c] = 0.f;
}

__device__ __forceinline__ float4 warp_layernorm4(float4 v, const float* __restrict__ gw, const float* __restrict__ gb, int lane) {
  float s = v.x + v.y + v.z + v.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / RENI_H);
  const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
  float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / RENI_H) + 1e-5f);
  const float4 w = *reinterpret_cast<const float4*>(gw + lane * 4), b = *reinterpret_cast<const float4*>(gb + lane * 4);
  return make_float4(d0 * rstd * w.x + b.x, d1 * rstd * w.y + b.y, d2 * rstd * w.z + b.z, d3 * rstd * w.w + b.w);
}

// x <- LN(x + add) * w + b, optionally followed in registers by x <- LN(x + add2) * w2 + b2 (norm2 of one decoder layer and norm1
// of the next are back to back: one pass over the activations instead of two)
__global__ void __launch_bounds__(256)
reni_ln_rows_kernel(float* __restrict__ x, int64_t N, const float* __restrict__ add, int add_stride, const int* __restrict__ row_cam,
                    const float* __restrict__ gw, const float* __restrict__ gb, const float* __restrict__ add2,
                    const float* __restrict__ gw2, const float* __restrict__ gb2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int64_t code = row_cam ? row_cam[row] : 0;
  float4 v = *reinterpret_cast<const float4*>(x + row * RENI_H + lane * 4);
  if (add) {
    const float4 a = *reinterpret_cast<const float4*>(add + code * add_stride + lane * 4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  v = warp_layernorm4(v, gw, gb, lane);
  if (gw2) {
    if (add2) {
      const float4 a = *reinterpret_cast<const float4*>(add2 + code * add_stride + lane * 4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    v = warp_layernorm4(v, gw2, gb2, lane);
  }
  *reinterpret_cast<float4*>(x + row * RENI_H + lane * 4) = v;
}

}  // namespace nsk

extern "C" int nsk_reni_prep(const float* latents, const float* rotation, int64_t K, const float* weights, int latent_dim, int hidden,
                             int num_layers, float* workspace, void* stream) {
  NSK_REQUIRE(hidden == nsk::RENI_H, "nsk_reni_prep: hidden_features must be 128");
  NSK_REQUIRE(latent_dim >= 1 && latent_dim <= nsk::RENI_MAX_L, "nsk_reni_prep: latent_dim out of range");
  if (K == 0) return 0;
  NSK_REQUIRE(latents && weights && workspace && K <= 65535, "nsk_reni_prep: null pointer / too many codes");
  const nsk::ReniLayout y = nsk::reni_layout(latent_dim, hidden, num_layers);
  float* attn = workspace;
  float* zxy = workspace + K * num_layers * (int64_t)hidden;
  return nsk::reni_launch_prep(latents, rotation, weights, y, K, zxy, attn, nsk::as_stream(stream));
}

extern "C" int nsk_reni_pe_rows(const float* dirs, const int* row_cam, int64_t N, const float* zxy, int latent_dim, float* pe, void* stream) {
  if (N == 0) return 0;
  NSK_REQUIRE(dirs && zxy && pe, "nsk_reni_pe_rows: null pointer");
  NSK_REQUIRE(latent_dim >= 1 && 5 * (latent_dim + 2) <= nsk::RENI_PE_LD, "nsk_reni_pe_rows: latent_dim out of range");
  const int64_t blocks = (N + 7) / 8;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_reni_pe_rows: too many rows for one launch");
  nsk::reni_pe_rows_kernel<<<(unsigned)blocks, 256, 0, nsk::as_stream(stream)>>>(dirs, row_cam, N, zxy, latent_dim, pe);
  return nsk::check_launch("reni_pe_rows_kernel");
}

extern "C" int nsk_reni_ln_rows(float* x, int64_t N, const float* add, int add_stride, const int* row_cam, const float* ln_weight,
                                const float* ln_bias, const float* add2, const float* ln_weight2, const float* ln_bias2, void* stream) {
  if (N == 0) return 0;
  NSK_REQUIRE(x && ln_weight && ln_bias && ((ln_weight2 == nullptr) == (ln_bias2 == nullptr)), "nsk_reni_ln_rows: null pointer");
  NSK_REQUIRE((add == nullptr && add2 == nullptr) || (add_stride & 3) == 0, "nsk_reni_ln_rows: add_stride must be a multiple of 4");
  const int64_t blocks = (N + 7) / 8;
  NSK_REQUIRE(blocks < (1ll << 31), "nsk_reni_ln_rows: too many rows for one launch");
  nsk::reni_ln_rows_kernel<<<(unsigned)blocks, 256, 0, nsk::as_stream(stream)>>>(x, N, add, add_stride, row_cam, ln_weight, ln_bias, add2, ln_weight2, ln_bias2);
  return nsk::check_launch("reni_ln_rows_kernel");
}

// K2 (exact fp32 CUDA-core path): SDF / albedo field of one sample, fused:
//   scene contraction -> hash-grid encode -> [x | PE6(x) | feat] -> geo MLP (softplus beta=100) -> sdf, geo feature
//   -> ANALYTIC d sdf / d x (reverse pass through the MLP, the PE and the trilinear hash interpolation)
//   -> colour MLP (ReLU, sigmoid) on [x | PE6(x) | geo feature].
// Replaces SDFAlbedoField.get_outputs / get_sdf_at_pos (neusky/fields/sdf_albedo_field.py:169-174, 211-269):
// forward_geonetwork [NS-mem, SURVEY A.4] (:233), torch.autograd.grad(sdf, x) (:235-238), get_colors (:185-209, :241).
// The reference keeps the autograd graph of the whole geo network alive to get the normals; here the gradient is a
// second, transposed pass over weights that are already on chip.
//
// This is the full-precision path (parity vs the fp32 oracle <= 1e-4); the throughput path is sdf_field_tc.cu.
// One CTA = 32 samples; thread n owns output column n of each 256-wide layer; activations live in shared memory as
// [k][row]; forward weights are pre-transposed to [K][N], backward weights keep torch's [out][in] layout, so both
// passes read weights coalesced.
#include "nsk_common.cuh"
#include "sdf_common.cuh"

namespace nsk {

constexpr int FR = 32;            // samples (rows) per CTA
constexpr int FT = SDF_HID;       // threads per CTA

struct SdfLayout {
  int64_t w0t, b0, w1t, b1, w2t, b2, w2s, b2s, w1, w0, c0t, c0b, c1t, c1b, c2, c2b, total;
};
__host__ __device__ inline SdfLayout sdf_layout() {
  SdfLayout y;
  int64_t o = 0;
  y.w0t = o; o += (int64_t)SDF_IN * SDF_HID;      // [71][256]
  y.b0 = o; o += SDF_HID;
  y.w1t = o; o += (int64_t)SDF_HID * SDF_HID;     // [256][256]
  y.b1 = o; o += SDF_HID;
  y.w2t = o; o += (int64_t)SDF_HID * SDF_HID;     // [256][256] geo-feature rows of the last layer (outputs 1..256)
  y.b2 = o; o += SDF_HID;
  y.w2s = o; o += SDF_HID;                        // row 0 of the last layer (sdf)
  y.b2s = o; o += 4;
  y.w1 = o; o += (int64_t)SDF_HID * SDF_HID;      // [out][in] for the reverse pass
  y.w0 = o; o += (int64_t)SDF_HID * SDF_IN;       // [out][in]
  y.c0t = o; o += (int64_t)SDF_CIN * SDF_HID;     // [295][256]
  y.c0b = o; o += SDF_HID;
  y.c1t = o; o += (int64_t)SDF_HID * SDF_HID;
  y.c1b = o; o += SDF_HID;
  y.c2 = o; o += 3 * SDF_HID;                     // [3][256]
  y.c2b = o; o += 4;
  y.total = o;
  return y;
}

__device__ __forceinline__ float softplus100(float z, float& sig) {
  // torch.nn.functional.softplus(beta=100, threshold=20): z if 100 z > 20 else log1p(exp(100 z)) / 100; sig = d/dz
  const float t = 100.0f * z;
  sig = 1.0f / (1.0f + expf(-t));
  return t > 20.0f ? z : log1pf(expf(t)) * 0.01f;
}

// acc[r] += sum_k in[k][r] * wt[k*ldw]
__device__ __forceinline__ void dense32(float (&acc)[FR], const float* __restrict__ in, const float* __restrict__ wt, int K, int ldw) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(wt + (int64_t)k * ldw);
    const float4* a = reinterpret_cast<const float4*>(in + k * FR);
#pragma unroll
    for (int v = 0; v < FR / 4; ++v) {
      const float4 x = a[v];
      acc[4 * v + 0] = fmaf(x.x, w, acc[4 * v + 0]);
      acc[4 * v + 1] = fmaf(x.y, w, acc[4 * v + 1]);
      acc[4 * v + 2] = fmaf(x.z, w, acc[4 * v + 2]);
      acc[4 * v + 3] = fmaf(x.w, w, acc[4 * v + 3]);
    }
  }
}

__global__ void __launch_bounds__(FT, 1)
sdf_field_simt_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ W, SdfLayout y,
                      const float2* __restrict__ table, const float* __restrict__ scalings, int log2_T, int flags,
                      float* __restrict__ sdf_out, float* __restrict__ grad_out, float* __restrict__ albedo_out,
                      float* __restrict__ geo_out, const GridMode gm) {
  extern __shared__ __align__(16) float smem[];
  float* in = smem;                         // [72][FR]   x, PE, hash features
  float* bufA = in + 72 * FR;               // [256][FR]  a0 -> geo feature
  float* bufB = bufA + SDF_HID * FR;        // [256][FR]  a1 -> c0
  float* bufS0 = bufB + SDF_HID * FR;       // [256][FR]  sigmoid(100 z0) -> g1 -> c1
  float* bufS1 = bufS0 + SDF_HID * FR;      // [256][FR]  sigmoid(100 z1) -> g2
  float* g0 = bufS1 + SDF_HID * FR;         // [72][FR]   d sdf / d input
  float* xs = g0 + 72 * FR;                 // [3][FR] raw x ; [3][FR] contracted pos ; [9][FR] Jacobian ; [3][FR] gx
  float* posb = xs + 3 * FR;
  float* jac = posb + 3 * FR;
  float* gx = jac + 9 * FR;
  const int t = threadIdx.x;
  const bool want_grad = flags & 1, want_colour = flags & 2;
  const uint32_t mask = (1u << log2_T) - 1u;
  const int64_t n_tiles = (n + FR - 1) / FR;
  const float TWO_PI = 6.283185307179586f, HALF_PI = 1.5707963267948966f;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * FR;
    // ---- inputs: x, contraction, PE -------------------------------------------------------------
    if (t < FR) {
      const int64_t r = min(row0 + t, n - 1);
      const float xv[3] = {x[r * 3], x[r * 3 + 1], x[r * 3 + 2]};
      float pos[3], J[9];
      sdf_contract(xv, pos, J);
#pragma unroll
      for (int d = 0; d < 3; ++d) { xs[d * FR + t] = xv[d]; posb[d * FR + t] = pos[d]; in[d * FR + t] = xv[d]; gx[d * FR + t] = 0.f; }
#pragma unroll
      for (int i = 0; i < 9; ++i) jac[i * FR + t] = J[i];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float s = TWO_PI * xv[d];
#pragma unroll
        for (int f = 0; f < 6; ++f) {
          const float a = s * (float)(1 << f);
          in[(3 + d * 6 + f) * FR + t] = sinf(a);
          in[(3 + 18 + d * 6 + f) * FR + t] = sinf(a + HALF_PI);
        }
      }
    }
    __syncthreads();
    // ---- hash features of the contracted position --------------------------------------------------
    {
      const int r = t % FR, sub = t / FR;
      const float px = posb[r], py = posb[FR + r], pz = posb[2 * FR + r];
      for (int lev = sub; lev < SDF_LEVELS; lev += FT / FR) {
        uint32_t idx[8];
        float ox, oy, oz, dw[3], s;
        grid_corners(gm, lev, px, py, pz, scalings[lev], mask, idx, ox, oy, oz, dw, s);
        const float2* tl = table + ((size_t)lev << log2_T);
        float2 f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
        const float2 e = hash_interp(f, ox, oy, oz);
        in[(39 + 2 * lev) * FR + r] = e.x;
        in[(39 + 2 * lev + 1) * FR + r] = e.y;
      }
    }
    __syncthreads();
    // ---- geo layer 0 -----------------------------------------------------------------------------
    {
      float acc[FR];
      const float b = W[y.b0 + t];
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r] = b;
      dense32(acc, in, W + y.w0t + t, SDF_IN, SDF_HID);
#pragma unroll
      for (int r = 0; r < FR; ++r) { float sg; bufA[t * FR + r] = softplus100(acc[r], sg); bufS0[t * FR + r] = sg; }
    }
    __syncthreads();
    // ---- geo layer 1 -----------------------------------------------------------------------------
    {
      float acc[FR];
      const float b = W[y.b1 + t];
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r] = b;
      dense32(acc, bufA, W + y.w1t + t, SDF_HID, SDF_HID);
      const float w2 = W[y.w2s + t];
#pragma unroll
      for (int r = 0; r < FR; ++r) { float sg; bufB[t * FR + r] = softplus100(acc[r], sg); bufS1[t * FR + r] = sg * w2; }   // g2 = W2[0,:] * s1
    }
    __syncthreads();
    // ---- geo layer 2: geo feature (outputs 1..256) and sdf (output 0) -------------------------------
    {
      float acc[FR];
      const float b = W[y.b2 + t];
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r] = b;
      dense32(acc, bufB, W + y.w2t + t, SDF_HID, SDF_HID);
#pragma unroll
      for (int r = 0; r < FR; ++r) bufA[t * FR + r] = acc[r];        // a0 is dead; geo feature lives here
      if (geo_out) {
#pragma unroll 4
        for (int r = 0; r < FR; ++r) if (row0 + r < n) geo_out[(row0 + r) * SDF_HID + t] = acc[r];
      }
      if (t < FR && row0 + t < n) {
        float o = W[y.b2s];
        for (int k = 0; k < SDF_HID; ++k) o = fmaf(bufB[k * FR + t], W[y.w2s + k], o);
        sdf_out[row0 + t] = o;
      }
    }
    __syncthreads();
    if (want_grad) {
      // ---- reverse pass: g1 = (W1^T g2) * s0 --------------------------------------------------------
      {
        float acc[FR];
#pragma unroll
        for (int r = 0; r < FR; ++r) acc[r] = 0.f;
        dense32(acc, bufS1, W + y.w1 + t, SDF_HID, SDF_HID);      // W1[k][t]: [out k][in t]
        __syncthreads();   // (bufS0 is only read/written per-thread-row below, but keep phases explicit)
#pragma unroll
        for (int r = 0; r < FR; ++r) bufS0[t * FR + r] *= acc[r];
      }
      __syncthreads();
      // ---- g0 = W0^T g1 (71 inputs) ---------------------------------------------------------------
      if (t < SDF_IN) {
        float acc[FR];
#pragma unroll
        for (int r = 0; r < FR; ++r) acc[r] = 0.f;
        dense32(acc, bufS0, W + y.w0 + t, SDF_HID, SDF_IN);       // W0[k][t]
#pragma unroll
        for (int r = 0; r < FR; ++r) g0[t * FR + r] = acc[r];
      }
      __syncthreads();
      // ---- chain to x: hash interpolation (per row x level), then PE and identity (per row) --------------
      {
        const int r = t % FR, sub = t / FR;
        const float px = posb[r], py = posb[FR + r], pz = posb[2 * FR + r];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;    // d sdf / d pos accumulated over this thread's levels
        for (int lev = sub; lev < SDF_LEVELS; lev += FT / FR) {
          uint32_t idx[8];
          float ox, oy, oz, dw[3], s;
          grid_corners(gm, lev, px, py, pz, scalings[lev], mask, idx, ox, oy, oz, dw, s);
          const float2* tl = table + ((size_t)lev << log2_T);
          float fa[8], fb[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) { const float2 v = __ldg(tl + idx[c]); fa[c] = v.x; fb[c] = v.y; }
          float da[3], db[3];
          hash_interp_grad(fa, ox, oy, oz, da);
          hash_interp_grad(fb, ox, oy, oz, db);
          const float ga = g0[(39 + 2 * lev) * FR + r], gb = g0[(39 + 2 * lev + 1) * FR + r];
          a0 += s * dw[0] * (da[0] * ga + db[0] * gb);
          a1 += s * dw[1] * (da[1] * ga + db[1] * gb);
          a2 += s * dw[2] * (da[2] * ga + db[2] * gb);
        }
        // d sdf / d x += J^T (d sdf / d pos)
        atomicAdd(&gx[0 * FR + r], jac[0 * FR + r] * a0 + jac[3 * FR + r] * a1 + jac[6 * FR + r] * a2);
        atomicAdd(&gx[1 * FR + r], jac[1 * FR + r] * a0 + jac[4 * FR + r] * a1 + jac[7 * FR + r] * a2);
        atomicAdd(&gx[2 * FR + r], jac[2 * FR + r] * a0 + jac[5 * FR + r] * a1 + jac[8 * FR + r] * a2);
      }
      __syncthreads();
      if (t < FR && row0 + t < n) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float g = gx[d * FR + t] + g0[d * FR + t];
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            const float k = TWO_PI * (float)(1 << f);
            const float sn = in[(3 + d * 6 + f) * FR + t], cs = in[(3 + 18 + d * 6 + f) * FR + t];
            g += k * (cs * g0[(3 + d * 6 + f) * FR + t] - sn * g0[(3 + 18 + d * 6 + f) * FR + t]);
          }
          grad_out[(row0 + t) * 3 + d] = g;
        }
      }
      __syncthreads();
    }
    if (want_colour) {
      // ---- colour layer 0: [x | PE | geo] -> 256, ReLU --------------------------------------------------
      {
        float acc[FR];
        const float b = W[y.c0b + t];
#pragma unroll
        for (int r = 0; r < FR; ++r) acc[r] = b;
        dense32(acc, in, W + y.c0t + t, 39, SDF_HID);
        dense32(acc, bufA, W + y.c0t + 39 * SDF_HID + t, SDF_HID, SDF_HID);
#pragma unroll
        for (int r = 0; r < FR; ++r) bufB[t * FR + r] = fmaxf(acc[r], 0.f);
      }
      __syncthreads();
      {
        float acc[FR];
        const float b = W[y.c1b + t];
#pragma unroll
        for (int r = 0; r < FR; ++r) acc[r] = b;
        dense32(acc, bufB, W + y.c1t + t, SDF_HID, SDF_HID);
#pragma unroll
        for (int r = 0; r < FR; ++r) bufS0[t * FR + r] = fmaxf(acc[r], 0.f);
      }
      __syncthreads();
      if (t < 3 * FR) {
        const int r = t % FR, c = t / FR;
        if (row0 + r < n) {
          float o = W[y.c2b + c];
          for (int k = 0; k < SDF_HID; ++k) o = fmaf(bufS0[k * FR + r], W[y.c2 + c * SDF_HID + k], o);
          albedo_out[(row0 + r) * 3 + c] = sigmoidf_(o);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace nsk

extern "C" int64_t nsk_sdf_simt_weights_floats(void) { return nsk::sdf_layout().total; }

extern "C" int nsk_sdf_field_simt_fwd(const float* x, int64_t n, const float* sdf_weights, const float* hash_table,
                                      const float* scalings, int num_levels, int log2_T, float* sdf, float* grad,
                                      float* albedo, float* geo, void* stream) {
  return nsk_sdf_field_simt_fwd_ex(x, n, sdf_weights, hash_table, scalings, num_levels, log2_T, nullptr, 0, sdf, grad, albedo, geo, stream);
}

extern "C" int nsk_sdf_field_simt_fwd_ex(const float* x, int64_t n, const float* sdf_weights, const float* hash_table,
                                         const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                                         float* sdf, float* grad, float* albedo, float* geo, void* stream) {
  NSK_REQUIRE(num_levels == nsk::SDF_LEVELS, "nsk_sdf_field_simt_fwd: the SDF position encoding has 16 levels");
  NSK_REQUIRE(grid_meta == nullptr || (reinterpret_cast<uintptr_t>(grid_meta) & 15) == 0, "nsk_sdf_field_simt_fwd_ex: grid_meta must be 16-byte aligned");
  if (n == 0) return 0;
  NSK_REQUIRE(x && sdf_weights && hash_table && scalings && sdf, "nsk_sdf_field_simt_fwd: null pointer");
  const size_t smem = (size_t)(72 + 4 * nsk::SDF_HID + 72 + 18) * nsk::FR * sizeof(float);
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "cudaFuncSetAttribute(sdf_field_simt_kernel)", &num_sms, [&] {
        return cudaFuncSetAttribute(nsk::sdf_field_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }))
    return err;
  const int64_t n_tiles = (n + nsk::FR - 1) / nsk::FR;
  const int64_t grid = n_tiles < num_sms ? n_tiles : num_sms;
  const int flags = (grad ? 1 : 0) | (albedo ? 2 : 0);
  nsk::sdf_field_simt_kernel<<<(unsigned)grid, nsk::FT, smem, nsk::as_stream(stream)>>>(
      x, n, sdf_weights, nsk::sdf_layout(), reinterpret_cast<const float2*>(hash_table), scalings, log2_T, flags, sdf, grad,
      albedo, geo, nsk::GridMode{reinterpret_cast<const int4*>(grid_meta), smoothstep});
  return nsk::check_launch("sdf_field_simt_kernel");
}

// K4 (exact fp32 CUDA-core path): DDF sky visibility fused with the cosine-weighted Lambertian sum.
// Replaces NeuSkyFactoModel.compute_visibility (neusky/models/neusky_model.py:1624-1778),
// DDFModel.get_outputs (neusky/models/ddf_model.py:158-219), DirectionalDistanceField.get_outputs
// (neusky/fields/directional_distance_field.py:261-306), FiLMSiren (ns_reni/reni/field_components/
// film_siren.py:45-156) and the visibility-weighted einsum of the Lambertian renderer
// (neusky/model_components/renderers.py:106-113) for every (ray, direction) pair.
//
// This is the full-precision reference path on the GPU (parity <= 1e-4 vs the fp32 oracle); the
// throughput path is sky_shade_tc.cu (tcgen05).  One CTA = 32 pairs; thread n owns output column n
// of each 256-wide layer; activations live in shared memory as [k][row] so a thread reads 4 rows
// per 16-byte load; weights are pre-transposed to [K][N] so warps read them coalesced.
// The [R, D'] visibility tensor is only written when the caller asks for it.
#include "nsk_common.cuh"

namespace nsk {

constexpr int SR = 32;          // pairs (rows) per CTA
constexpr int ST = DDF_HID;     // threads per CTA == layer width

struct SimtLayout {
  int64_t map_wt[DDF_LAYERS + 1], map_b[DDF_LAYERS + 1];  // mapping layers 0..4 (35|256 -> 256), 5 (256 -> 2560)
  int64_t net_wt[DDF_LAYERS], net_b[DDF_LAYERS];          // trunk layers (15|256 -> 256)
  int64_t fin_w, fin_b, total;
};

__host__ __device__ inline SimtLayout simt_layout() {
  SimtLayout y;
  int64_t o = 0;
  for (int i = 0; i <= DDF_LAYERS; ++i) {
    const int K = (i == 0) ? DDF_MAP_IN : DDF_HID;
    const int N = (i == DDF_LAYERS) ? DDF_FILM : DDF_HID;
    y.map_wt[i] = o; o += (int64_t)K * N;
    y.map_b[i] = o; o += N;
  }
  for (int l = 0; l < DDF_LAYERS; ++l) {
    const int K = (l == 0) ? DDF_DIR_IN : DDF_HID;
    y.net_wt[l] = o; o += (int64_t)K * DDF_HID;
    y.net_b[l] = o; o += DDF_HID;
  }
  y.fin_w = o; o += DDF_HID;
  y.fin_b = o; o += 4;
  y.total = o;
  return y;
}

// acc[r] += sum_k in[k][r] * wt[k*ldw + n]
__device__ __forceinline__ void dense_cols(float (&acc)[SR], const float* __restrict__ in, const float* __restrict__ wt,
                                           int K, int ldw) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(wt + (int64_t)k * ldw);
    const float4* a = reinterpret_cast<const float4*>(in + k * SR);
#pragma unroll
    for (int v = 0; v < SR / 4; ++v) {
      const float4 x = a[v];
      acc[4 * v + 0] = fmaf(x.x, w, acc[4 * v + 0]);
      acc[4 * v + 1] = fmaf(x.y, w, acc[4 * v + 1]);
      acc[4 * v + 2] = fmaf(x.z, w, acc[4 * v + 2]);
      acc[4 * v + 3] = fmaf(x.w, w, acc[4 * v + 3]);
    }
  }
}

__global__ void __launch_bounds__(ST, 2)
sky_shade_simt_kernel(const float* __restrict__ points, int64_t R, const float* __restrict__ normals,
                      const float* __restrict__ wa, const float* __restrict__ inv_count, int S,
                      const float* __restrict__ dirs, int Dp, const float* __restrict__ radiance,
                      const int32_t* __restrict__ cam, const float* __restrict__ W, SimtLayout y,
                      const float2* __restrict__ table, const float* __restrict__ scalings, int L, int log2_T,
                      float radius, float thr, float sig_scale, float* __restrict__ rgb_lin, float* __restrict__ vis_out,
                      float* __restrict__ ddf_out, float* __restrict__ term_out, const GridMode gm) {
  extern __shared__ __align__(16) float smem[];
  float* bufM = smem;                        // [256][SR]  mapping activations (m5 at the end)
  float* bufA = bufM + DDF_HID * SR;         // [256][SR]
  float* bufB = bufA + DDF_HID * SR;         // [256][SR]
  float* inM = bufB + DDF_HID * SR;          // [36][SR]   mapping input (q, hash features)
  float* inH = inM + 36 * SR;                // [16][SR]   trunk input (d_local, PE)
  float* geo = inH + 16 * SR;                // [4][SR]    q.xyz, term_dist
  const int t = threadIdx.x;
  const int64_t n_pairs = R * (int64_t)Dp;
  const int64_t n_tiles = (n_pairs + SR - 1) / SR;
  const uint32_t mask = (1u << log2_T) - 1u;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pair0 = tile * SR;
    // ---- prologue: pair geometry -------------------------------------------------------------
    if (t < SR) {
      const int64_t pr = min(pair0 + t, n_pairs - 1);
      const int64_t ray = pr / Dp;
      const int j = (int)(pr % Dp);
      const float p[3] = {points[ray * 3], points[ray * 3 + 1], points[ray * 3 + 2]};
      const float l[3] = {dirs[j * 3], dirs[j * 3 + 1], dirs[j * 3 + 2]};
      float q[3], tt;
      sphere_exit(p, l, radius, q, tt);                      // neusky_model.py:1693
      const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
      const float term = sqrtf(dx * dx + dy * dy + dz * dz);  // neusky_model.py:1697
      const float dneg[3] = {-l[0], -l[1], -l[2]};             // neusky_model.py:1702
      float dl[3], feat[15];
      ddf_local_dir(q, dneg, dl);
      ddf_dir_features(dl, feat);
#pragma unroll
      for (int c = 0; c < 15; ++c) inH[c * SR + t] = feat[c];
      geo[0 * SR + t] = q[0]; geo[1 * SR + t] = q[1]; geo[2 * SR + t] = q[2]; geo[3 * SR + t] = term;
      inM[0 * SR + t] = q[0]; inM[1 * SR + t] = q[1]; inM[2 * SR + t] = q[2];
    }
    __syncthreads();
    // ---- prologue: hash features of q (directional_distance_field.py:268) ----------------------
    {
      const int r = t % SR, sub = t / SR;  // 8 sub-groups x 2 levels
      const float qx = geo[r], qy = geo[SR + r], qz = geo[2 * SR + r];
      for (int lev = sub; lev < L; lev += ST / SR) {
        uint32_t idx[8];
        float ox, oy, oz, dw_[3], sc_;
        grid_corners(gm, lev, qx, qy, qz, scalings[lev], mask, idx, ox, oy, oz, dw_, sc_);
        const float2* tl = table + ((size_t)lev << log2_T);
        float2 f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
        const float2 e = hash_interp(f, ox, oy, oz);
        inM[(3 + 2 * lev) * SR + r] = e.x;
        inM[(3 + 2 * lev + 1) * SR + r] = e.y;
      }
    }
    __syncthreads();
    // ---- mapping network: 35 -> 256 x5, LeakyReLU(0.2)  (film_siren.py:45-58) -------------------
    {
      const float* src = inM;
      float* dst = bufM;
      for (int i = 0; i < DDF_LAYERS; ++i) {
        float acc[SR];
        const float b = W[y.map_b[i] + t];
#pragma unroll
        for (int r = 0; r < SR; ++r) acc[r] = b;
        dense_cols(acc, src, W + y.map_wt[i] + t, i == 0 ? DDF_MAP_IN : DDF_HID, DDF_HID);
#pragma unroll
        for (int r = 0; r < SR; ++r) dst[t * SR + r] = acc[r] > 0.f ? acc[r] : 0.2f * acc[r];
        __syncthreads();
        src = dst;
        dst = (dst == bufM) ? bufA : bufM;  // in->M, M->A, A->M, M->A, A->M : ends in bufM
      }
    }
    // ---- FiLM-SIREN trunk (film_siren.py:138-147) --------------------------------------------------
    {
      const float* hin = inH;
      float* hout = bufA;
      for (int l = 0; l < DDF_LAYERS; ++l) {
        float z[SR], g[SR];
        {
          const float b = W[y.net_b[l] + t];
#pragma unroll
          for (int r = 0; r < SR; ++r) z[r] = b;
          dense_cols(z, hin, W + y.net_wt[l] + t, l == 0 ? DDF_DIR_IN : DDF_HID, DDF_HID);
        }
        {
          const float b = W[y.map_b[DDF_LAYERS] + l * DDF_HID + t];
#pragma unroll
          for (int r = 0; r < SR; ++r) g[r] = b;
          dense_cols(g, bufM, W + y.map_wt[DDF_LAYERS] + l * DDF_HID + t, DDF_HID, DDF_FILM);
#pragma unroll
          for (int r = 0; r < SR; ++r) z[r] = (g[r] * 15.0f + 30.0f) * z[r];   // freq * x  (:140, :81)
        }
        {
          const float b = W[y.map_b[DDF_LAYERS] + DDF_FILM / 2 + l * DDF_HID + t];
#pragma unroll
          for (int r = 0; r < SR; ++r) g[r] = b;
          dense_cols(g, bufM, W + y.map_wt[DDF_LAYERS] + DDF_FILM / 2 + l * DDF_HID + t, DDF_HID, DDF_FILM);
#pragma unroll
          for (int r = 0; r < SR; ++r) hout[t * SR + r] = sinf(z[r] + g[r]);      // sin(freq*x + phase)
        }
        __syncthreads();
        hin = hout;
        hout = (hout == bufA) ? bufB : bufA;
      }
      // ---- final linear + sigmoid, visibility, shading (threads 0..31 = rows) --------------------
      if (t < SR && pair0 + t < n_pairs) {
        float o = W[y.fin_b];
        for (int k = 0; k < DDF_HID; ++k) o = fmaf(hin[k * SR + t], W[y.fin_w + k], o);
        const float ddf = sigmoidf_(o) * (2.0f * radius);       // directional_distance_field.py:297-299
        const float term = geo[3 * SR + t];
        const float vis = visibility_from_ddf(ddf, term, radius, thr, sig_scale);
        const int64_t pr = pair0 + t;
        const int64_t ray = pr / Dp;
        const int j = (int)(pr % Dp);
        if (vis_out) vis_out[pr] = vis;
        if (ddf_out) ddf_out[pr] = ddf;
        if (term_out) term_out[pr] = term;
        const float lx = dirs[j * 3], ly = dirs[j * 3 + 1], lz = dirs[j * 3 + 2];
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int s = 0; s < S; ++s) {
          const int64_t i = ray * S + s;
          float c = normals[i * 3] * lx + normals[i * 3 + 1] * ly + normals[i * 3 + 2] * lz;
          c = fminf(fmaxf(c, 0.f), 1.f) * inv_count[i];
          c0 += wa[i * 3] * c; c1 += wa[i * 3 + 1] * c; c2 += wa[i * 3 + 2] * c;
        }
        const float* rad = radiance + ((int64_t)(cam ? cam[ray] : 0) * Dp + j) * 3;
        atomicAdd(rgb_lin + ray * 3 + 0, c0 * vis * rad[0]);
        atomicAdd(rgb_lin + ray * 3 + 1, c1 * vis * rad[1]);
        atomicAdd(rgb_lin + ray * 3 + 2, c2 * vis * rad[2]);
      }
    }
    __syncthreads();
  }
}

}  // namespace nsk

extern "C" int64_t nsk_ddf_simt_weights_floats(void) { return nsk::simt_layout().total; }

extern "C" int nsk_sky_shade_simt_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                                      const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                                      const int32_t* cam, const float* ddf_weights, const float* hash_table,
                                      const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                                      float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out,
                                      float* term_out, void* stream) {
  return nsk_sky_shade_simt_fwd_ex(points, R, normals, wa, inv_count, S, dirs, Dp, radiance, cam, ddf_weights, hash_table, scalings, num_levels, log2_T,
                                   nullptr, 0, radius, threshold, sigmoid_scale, rgb_lin, vis_out, ddf_out, term_out, stream);
}

extern "C" int nsk_sky_shade_simt_fwd_ex(const float* points, int64_t R, const float* normals, const float* wa,
                                         const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                                         const int32_t* cam, const float* ddf_weights, const float* hash_table,
                                         const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                                         float radius, float threshold, float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out,
                                         float* term_out, void* stream) {
  NSK_REQUIRE(grid_meta == nullptr || (reinterpret_cast<uintptr_t>(grid_meta) & 15) == 0, "nsk_sky_shade_simt_fwd_ex: grid_meta must be 16-byte aligned");
  NSK_REQUIRE(num_levels == nsk::DDF_LEVELS, "nsk_sky_shade_simt_fwd: the DDF position encoding has 16 levels");
  if (R == 0 || Dp == 0) return 0;
  NSK_REQUIRE(S >= 1, "nsk_sky_shade_simt_fwd: S must be >= 1");
  NSK_REQUIRE(points && normals && wa && inv_count && dirs && radiance && ddf_weights && hash_table && scalings && rgb_lin,
              "nsk_sky_shade_simt_fwd: null pointer");
  const size_t smem = (size_t)(3 * nsk::DDF_HID + 36 + 16 + 4) * nsk::SR * sizeof(float);
  static nsk::DeviceOnce once;
  int num_sms = 0;
  if (int err = nsk::device_once(once, "cudaFuncSetAttribute(sky_shade_simt_kernel)", &num_sms, [&] {
        return cudaFuncSetAttribute(nsk::sky_shade_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }))
    return err;
  const int64_t n_tiles = (R * (int64_t)Dp + nsk::SR - 1) / nsk::SR;
  const int64_t grid = n_tiles < num_sms * 2 ? n_tiles : num_sms * 2;
  nsk::sky_shade_simt_kernel<<<(unsigned)grid, nsk::ST, smem, nsk::as_stream(stream)>>>(
      points, R, normals, wa, inv_count, S, dirs, Dp, radiance, cam, ddf_weights, nsk::simt_layout(),
      reinterpret_cast<const float2*>(hash_table), scalings, num_levels, log2_T, radius, threshold, sigmoid_scale, rgb_lin,
      vis_out, ddf_out, term_out, nsk::GridMode{reinterpret_cast<const int4*>(grid_meta), smoothstep});
  return nsk::check_launch("sky_shade_simt_kernel");
}

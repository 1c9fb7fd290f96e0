// Training-path companions of the tf32 layer contractions (gemm_tf32.cu): everything between two GEMMs of the DDF
// visibility network in the training step, forward and backward.  Row r of every [N, .] tensor is one
// (surface point, light direction) pair, N = R * D' (neusky/models/neusky_model.py:1685-1690).
//
//   nsk_ddf_pairs_fwd   pair geometry: sphere exit q (neusky_model.py:1693-1695), local-frame direction and its NeRF
//                       encoding (neusky/models/ddf_model.py:158-200, directional_distance_field.py:270-271), hash-grid
//                       conditioning (directional_distance_field.py:267-268), clamped ground-truth distance (:1724-1727)
//   nsk_film_sin_fwd    FiLM layer activation sin((15 f + 30) z + phase) (film_siren.py:74-81, 140)
//   nsk_film_sin_bwd    its backward: d z, and d f / d phase written into the [N, 2560] mapping-output gradient
//   nsk_ddf_head_fwd    final 256 -> 1 layer, sigmoid * 2r (directional_distance_field.py:297-299), visibility sigmoid
//                       (neusky_model.py:1730-1740)
//   nsk_ddf_head_bwd    backward of the head: d a5, d w_final, d b_final, d threshold
//   nsk_colsum          bias gradients: out[c] += sum_r X[r, c]
// All fp32, memory-bound elementwise / row-reduction kernels; one pass over their operands each.
#include <algorithm>

#include "nsk_common.cuh"
#include "sdf_common.cuh"

namespace nsk {
namespace train {

constexpr int COND_LD = 40;   // 3 + 32 padded to a multiple of 8 (tf32 MMA K step)
constexpr int XIN_LD = 16;    // 15 padded

// cond = q | hash(q) (directional_distance_field.py:267-268), xin = d_local | PE2(d_local) (:270-271) for one row
__device__ __forceinline__ void ddf_row_inputs(const float q[3], const float d[3], const float2* __restrict__ table,
                                               const float* __restrict__ scalings, int L, int log2_T,
                                               float* __restrict__ cr, float* __restrict__ xr) {
  const uint32_t mask = (1u << log2_T) - 1u;
  float dl[3], feat[15];
  ddf_local_dir(q, d, dl);
  ddf_dir_features(dl, feat);
#pragma unroll
  for (int c = 0; c < 15; ++c) xr[c] = feat[c];
  xr[15] = 0.f;
  cr[0] = q[0]; cr[1] = q[1]; cr[2] = q[2];
  for (int lev = 0; lev < L; ++lev) {
    const float s = scalings[lev];
    uint32_t idx[8];
    float ox, oy, oz;
    hash_corners(__fmul_rn(q[0], s), __fmul_rn(q[1], s), __fmul_rn(q[2], s), mask, idx, ox, oy, oz);
    const float2* tl = table + ((size_t)lev << log2_T);
    float2 f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
    const float2 v = hash_interp(f, ox, oy, oz);
    cr[3 + 2 * lev] = v.x;
    cr[4 + 2 * lev] = v.y;
  }
  for (int c = 3 + 2 * L; c < COND_LD; ++c) cr[c] = 0.f;
}

__global__ void __launch_bounds__(128)
ddf_pairs_fwd_kernel(const float* __restrict__ points, int64_t R, const float* __restrict__ dirs, int D,
                     const float2* __restrict__ table, const float* __restrict__ scalings, int L, int log2_T, float radius,
                     float* __restrict__ cond, float* __restrict__ xin, float* __restrict__ qout, float* __restrict__ gt) {
  const int64_t N = R * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int j = (int)(i - r * D);
    const float p[3] = {points[r * 3], points[r * 3 + 1], points[r * 3 + 2]};
    const float l[3] = {dirs[j * 3], dirs[j * 3 + 1], dirs[j * 3 + 2]};
    float q[3], t;
    sphere_exit(p, l, radius, q, t);
    // termination_dist / dist_to_ray_origins: |q - p| (neusky_model.py:1697-1699, 1724)
    const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    gt[i] = dist;
    qout[i * 3] = q[0]; qout[i * 3 + 1] = q[1]; qout[i * 3 + 2] = q[2];
    const float dneg[3] = {-l[0], -l[1], -l[2]};
    ddf_row_inputs(q, dneg, table, scalings, L, log2_T, cond + i * COND_LD, xin + i * XIN_LD);
  }
}

// Same network inputs for rows that already carry their own sphere point and world direction: the DDF fitting pass
// (neusky/models/ddf_model.py:193-217 -- DDFModel.get_outputs on a sampled ray bundle, the multi-view and the sky-ray
// batches :279-363).
__global__ void __launch_bounds__(128)
ddf_rows_fwd_kernel(const float* __restrict__ origins, const float* __restrict__ directions, int64_t N,
                    const float2* __restrict__ table, const float* __restrict__ scalings, int L, int log2_T,
                    float* __restrict__ cond, float* __restrict__ xin) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float q[3] = {origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]};
    const float d[3] = {directions[i * 3], directions[i * 3 + 1], directions[i * 3 + 2]};
    ddf_row_inputs(q, d, table, scalings, L, log2_T, cond + i * COND_LD, xin + i * XIN_LD);
  }
}

// a = sin((15 F[:, l*H + c] + 30) * z + F[:, half + l*H + c]);  H = 256, F row stride ldf, half = ldf / 2
__global__ void __launch_bounds__(256)
film_sin_fwd_kernel(const float* __restrict__ z, const float* __restrict__ F, int ldf, int layer, int64_t N, float* __restrict__ a) {
  const int64_t total = N * 64;  // float4 granules of a [N,256] tensor
  const int half = ldf >> 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 6;
    const int c = (int)(i & 63) * 4;
    const float4 zz = *reinterpret_cast<const float4*>(z + r * 256 + c);
    const float4 fr = *reinterpret_cast<const float4*>(F + r * ldf + layer * 256 + c);
    const float4 ph = *reinterpret_cast<const float4*>(F + r * ldf + half + layer * 256 + c);
    float4 o;
    o.x = sinf(fmaf(fmaf(15.f, fr.x, 30.f), zz.x, ph.x));
    o.y = sinf(fmaf(fmaf(15.f, fr.y, 30.f), zz.y, ph.y));
    o.z = sinf(fmaf(fmaf(15.f, fr.z, 30.f), zz.z, ph.z));
    o.w = sinf(fmaf(fmaf(15.f, fr.w, 30.f), zz.w, ph.w));
    *reinterpret_cast<float4*>(a + r * 256 + c) = o;
  }
}

// du = da * cos(u); dz = du * freq; dF[:, l*H + c] = 15 * du * z; dF[:, half + l*H + c] = du
// SUMS: also accumulate the column sums of dz (the trunk layer's bias gradient) and of the two dF column blocks (their share of the
// last mapping layer's bias gradient) -- a thread keeps its 4 columns for the whole grid-stride loop (the grid is a multiple of 64
// granules), so the sums ride in registers and the separate colsum passes over [N,256] x 5 and [N,2560] go away.
template <bool SUMS>
__global__ void __launch_bounds__(256)
film_sin_bwd_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ F, int ldf, int layer,
                    int64_t N, float* __restrict__ dz, float* __restrict__ dF, float* __restrict__ sum_dz, float* __restrict__ sum_dF) {
  const int64_t total = N * 64;
  const int half = ldf >> 1;
  float4 s_z = make_float4(0.f, 0.f, 0.f, 0.f), s_f = s_z, s_p = s_z;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 6;
    const int c = (int)(i & 63) * 4;
    const float4 g = *reinterpret_cast<const float4*>(da + r * 256 + c);
    const float4 zz = *reinterpret_cast<const float4*>(z + r * 256 + c);
    const float4 fr = *reinterpret_cast<const float4*>(F + r * ldf + layer * 256 + c);
    const float4 ph = *reinterpret_cast<const float4*>(F + r * ldf + half + layer * 256 + c);
    float4 odz, odf, odp;
#define NSK_FILM_BWD(k)                                          \
  {                                                              \
    const float fq = fmaf(15.f, fr.k, 30.f);                     \
    const float du = g.k * cosf(fmaf(fq, zz.k, ph.k));           \
    odz.k = du * fq;                                             \
    odf.k = 15.f * du * zz.k;                                    \
    odp.k = du;                                                  \
  }
    NSK_FILM_BWD(x) NSK_FILM_BWD(y) NSK_FILM_BWD(z) NSK_FILM_BWD(w)
#undef NSK_FILM_BWD
    *reinterpret_cast<float4*>(dz + r * 256 + c) = odz;
    *reinterpret_cast<float4*>(dF + r * ldf + layer * 256 + c) = odf;
    *reinterpret_cast<float4*>(dF + r * ldf + half + layer * 256 + c) = odp;
    if (SUMS) {
      s_z.x += odz.x; s_z.y += odz.y; s_z.z += odz.z; s_z.w += odz.w;
      s_f.x += odf.x; s_f.y += odf.y; s_f.z += odf.z; s_f.w += odf.w;
      s_p.x += odp.x; s_p.y += odp.y; s_p.z += odp.z; s_p.w += odp.w;
    }
  }
  if (SUMS) {
    // 256 threads = 4 row groups x 64 column granules: fold the row groups through shared memory, one atomic per column and block
    __shared__ float4 sh[3][4][64];
    const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
    sh[0][rg][cg] = s_z; sh[1][rg][cg] = s_f; sh[2][rg][cg] = s_p;
    __syncthreads();
    if (threadIdx.x < 192) {
      const int which = threadIdx.x >> 6;
      float4 t = sh[which][0][cg];
#pragma unroll
      for (int g2 = 1; g2 < 4; ++g2) { const float4 u = sh[which][g2][cg]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
      float* dst = which == 0 ? sum_dz + cg * 4 : which == 1 ? sum_dF + layer * 256 + cg * 4 : sum_dF + half + layer * 256 + cg * 4;
      atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
    }
  }
}

// one warp per pair: o = a5 . w + b; that = 2 r sigmoid(o); vis = 1 - sigmoid(scale * (min(gt, 2r) - that - thr))
__global__ void __launch_bounds__(256)
ddf_head_fwd_kernel(const float* __restrict__ a5, const float* __restrict__ wf, const float* __restrict__ bf,
                    const float* __restrict__ gt, int64_t N, float radius, const float* __restrict__ thr_p, float scale,
                    float* __restrict__ that, float* __restrict__ vis) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4 w0 = *reinterpret_cast<const float4*>(wf + lane * 8), w1 = *reinterpret_cast<const float4*>(wf + lane * 8 + 4);
  const float b = bf[0], thr = thr_p[0];
  for (int64_t r = warp; r < N; r += nwarps) {
    const float4 x0 = *reinterpret_cast<const float4*>(a5 + r * 256 + lane * 8), x1 = *reinterpret_cast<const float4*>(a5 + r * 256 + lane * 8 + 4);
    float s = x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
    s = warp_sum(s);
    if (lane == 0) {
      const float t = 2.0f * radius * sigmoidf_(s + b);
      that[r] = t;
      vis[r] = visibility_from_ddf(t, gt[r], radius, thr, scale);
    }
  }
}

// d that = d vis * scale * occ (1 - occ) + d that_extra;  d o = d that * 2r * sig (1 - sig), sig = that / 2r
// d a5 = d o * w;  d w += sum d o * a5;  d b += sum d o;  d thr += sum d vis * scale * occ (1 - occ)
__global__ void __launch_bounds__(256)
ddf_head_bwd_kernel(const float* __restrict__ a5, const float* __restrict__ wf, const float* __restrict__ that,
                    const float* __restrict__ gt, const float* __restrict__ d_vis, const float* __restrict__ d_that_extra,
                    int64_t N, float radius, const float* __restrict__ thr_p, float scale, float* __restrict__ da5,
                    float* __restrict__ d_wf, float* __restrict__ d_bf, float* __restrict__ d_thr) {
  __shared__ float s_dw[8][256];
  __shared__ float s_db[8], s_dt[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4 w0 = *reinterpret_cast<const float4*>(wf + lane * 8), w1 = *reinterpret_cast<const float4*>(wf + lane * 8 + 4);
  const float thr = thr_p[0];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float acc_b = 0.f, acc_t = 0.f;
  for (int64_t r = warp; r < N; r += nwarps) {
    const float t = that[r];
    const float g = fminf(gt[r], 2.0f * radius);
    const float occ = sigmoidf_(scale * ((g - t) - thr));
    const float dv = d_vis ? d_vis[r] : 0.f;
    const float k = dv * scale * occ * (1.0f - occ);
    float dt = k + (d_that_extra ? d_that_extra[r] : 0.f);
    const float sg = t / (2.0f * radius);
    const float d_o = dt * 2.0f * radius * sg * (1.0f - sg);
    const float4 x0 = *reinterpret_cast<const float4*>(a5 + r * 256 + lane * 8), x1 = *reinterpret_cast<const float4*>(a5 + r * 256 + lane * 8 + 4);
    *reinterpret_cast<float4*>(da5 + r * 256 + lane * 8) = make_float4(d_o * w0.x, d_o * w0.y, d_o * w0.z, d_o * w0.w);
    *reinterpret_cast<float4*>(da5 + r * 256 + lane * 8 + 4) = make_float4(d_o * w1.x, d_o * w1.y, d_o * w1.z, d_o * w1.w);
    acc[0] = fmaf(d_o, x0.x, acc[0]); acc[1] = fmaf(d_o, x0.y, acc[1]); acc[2] = fmaf(d_o, x0.z, acc[2]); acc[3] = fmaf(d_o, x0.w, acc[3]);
    acc[4] = fmaf(d_o, x1.x, acc[4]); acc[5] = fmaf(d_o, x1.y, acc[5]); acc[6] = fmaf(d_o, x1.z, acc[6]); acc[7] = fmaf(d_o, x1.w, acc[7]);
    acc_b += d_o;
    acc_t += k;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) s_dw[wib][lane * 8 + c] = acc[c];
  if (lane == 0) { s_db[wib] = acc_b; s_dt[wib] = acc_t; }
  __syncthreads();
  {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_dw[w][threadIdx.x];
    atomicAdd(d_wf + threadIdx.x, s);
  }
  if (threadIdx.x == 0) {
    float sb = 0.f, stt = 0.f;
    for (int w = 0; w < 8; ++w) { sb += s_db[w]; stt += s_dt[w]; }
    atomicAdd(d_bf, sb);
    if (d_thr) atomicAdd(d_thr, stt);
  }
}

// out[c] += sum_r X[r, c] ; block = 256 threads owning 256 consecutive columns of a row slab
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int ld, int64_t M, int ncols, int64_t rows_per_block, float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (c >= ncols) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += X[r * ld + c];
    s1 += X[(r + 1) * ld + c];
    s2 += X[(r + 2) * ld + c];
    s3 += X[(r + 3) * ld + c];
  }
  for (; r < r1; ++r) s0 += X[r * ld + c];
  atomicAdd(out + c, (s0 + s1) + (s2 + s3));
}

static int grid_for(int64_t work_items, int block, int max_blocks = 148 * 8) {
  const int64_t b = (work_items + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, max_blocks));
}

}  // namespace train
}  // namespace nsk

using namespace nsk;
using namespace nsk::train;

extern "C" int nsk_ddf_pairs_fwd(const float* points, int64_t R, const float* dirs, int D, const float* table,
                                 const float* scalings, int num_levels, int log2_T, float radius, float* cond, float* xin,
                                 float* q, float* gt, void* stream) {
  NSK_REQUIRE(points && dirs && table && scalings && cond && xin && q && gt, "nsk_ddf_pairs_fwd: null pointer");
  NSK_REQUIRE(num_levels == 16 && log2_T > 0 && log2_T < 31, "nsk_ddf_pairs_fwd: hash grid shape");
  if (R * D == 0) return 0;
  ddf_pairs_fwd_kernel<<<grid_for(R * D, 128, 148 * 16), 128, 0, as_stream(stream)>>>(
      points, R, dirs, D, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, radius, cond, xin, q, gt);
  return check_launch("ddf_pairs_fwd_kernel");
}

extern "C" int nsk_ddf_rows_fwd(const float* origins, const float* directions, int64_t N, const float* table,
                                const float* scalings, int num_levels, int log2_T, float* cond, float* xin, void* stream) {
  NSK_REQUIRE(origins && directions && table && scalings && cond && xin, "nsk_ddf_rows_fwd: null pointer");
  NSK_REQUIRE(num_levels == 16 && log2_T > 0 && log2_T < 31, "nsk_ddf_rows_fwd: hash grid shape");
  if (N == 0) return 0;
  ddf_rows_fwd_kernel<<<grid_for(N, 128, 148 * 16), 128, 0, as_stream(stream)>>>(
      origins, directions, N, reinterpret_cast<const float2*>(table), scalings, num_levels, log2_T, cond, xin);
  return check_launch("ddf_rows_fwd_kernel");
}

extern "C" int nsk_film_sin_fwd(const float* z, const float* film, int ldf, int layer, int64_t N, float* a, void* stream) {
  NSK_REQUIRE(z && film && a, "nsk_film_sin_fwd: null pointer");
  NSK_REQUIRE((ldf & 7) == 0 && layer >= 0 && (layer + 1) * 256 <= ldf / 2, "nsk_film_sin_fwd: layer / ldf");
  if (N == 0) return 0;
  film_sin_fwd_kernel<<<grid_for(N * 64, 256), 256, 0, as_stream(stream)>>>(z, film, ldf, layer, N, a);
  return check_launch("film_sin_fwd_kernel");
}

extern "C" int nsk_film_sin_bwd(const float* da, const float* z, const float* film, int ldf, int layer, int64_t N, float* dz,
                                float* dfilm, void* stream) {
  NSK_REQUIRE(da && z && film && dz && dfilm, "nsk_film_sin_bwd: null pointer");
  NSK_REQUIRE((ldf & 7) == 0 && layer >= 0 && (layer + 1) * 256 <= ldf / 2, "nsk_film_sin_bwd: layer / ldf");
  if (N == 0) return 0;
  film_sin_bwd_kernel<false><<<grid_for(N * 64, 256), 256, 0, as_stream(stream)>>>(da, z, film, ldf, layer, N, dz, dfilm, nullptr, nullptr);
  return check_launch("film_sin_bwd_kernel");
}

extern "C" int nsk_film_sin_bwd_sums(const float* da, const float* z, const float* film, int ldf, int layer, int64_t N, float* dz,
                                     float* dfilm, float* sum_dz, float* sum_dfilm, void* stream) {
  NSK_REQUIRE(da && z && film && dz && dfilm && sum_dz && sum_dfilm, "nsk_film_sin_bwd_sums: null pointer");
  NSK_REQUIRE((ldf & 7) == 0 && layer >= 0 && (layer + 1) * 256 <= ldf / 2, "nsk_film_sin_bwd_sums: layer / ldf");
  if (N == 0) return 0;
  // block = 256 threads = 4 rows of 64 granules, so every thread keeps its column granule across the grid-stride loop for any grid size
  film_sin_bwd_kernel<true><<<grid_for(N * 64, 256), 256, 0, as_stream(stream)>>>(da, z, film, ldf, layer, N, dz, dfilm, sum_dz, sum_dfilm);
  return check_launch("film_sin_bwd_kernel");
}

extern "C" int nsk_ddf_head_fwd(const float* a5, const float* w_final, const float* b_final, const float* gt, int64_t N,
                                float radius, const float* threshold, float sigmoid_scale, float* that, float* vis,
                                void* stream) {
  NSK_REQUIRE(a5 && w_final && b_final && gt && threshold && that && vis, "nsk_ddf_head_fwd: null pointer");
  if (N == 0) return 0;
  ddf_head_fwd_kernel<<<grid_for(N * 32, 256), 256, 0, as_stream(stream)>>>(a5, w_final, b_final, gt, N, radius, threshold,
                                                                         sigmoid_scale, that, vis);
  return check_launch("ddf_head_fwd_kernel");
}

extern "C" int nsk_ddf_head_bwd(const float* a5, const float* w_final, const float* that, const float* gt, const float* d_vis,
                                const float* d_that_extra, int64_t N, float radius, const float* threshold,
                                float sigmoid_scale, float* da5, float* d_w_final, float* d_b_final, float* d_threshold,
                                void* stream) {
  NSK_REQUIRE(a5 && w_final && that && gt && threshold && da5 && d_w_final && d_b_final, "nsk_ddf_head_bwd: null pointer");
  if (N == 0) return 0;
  ddf_head_bwd_kernel<<<grid_for(N * 32, 256, 148 * 4), 256, 0, as_stream(stream)>>>(
      a5, w_final, that, gt, d_vis, d_that_extra, N, radius, threshold, sigmoid_scale, da5, d_w_final, d_b_final, d_threshold);
  return check_launch("ddf_head_bwd_kernel");
}

extern "C" int nsk_colsum(const float* X, int ld, int64_t M, int ncols, float* out, void* stream) {
  NSK_REQUIRE(X && out && ld >= ncols && ncols > 0, "nsk_colsum: arguments");
  if (M == 0) return 0;
  const int cb = (ncols + 255) / 256;
  int64_t rb = std::max<int64_t>(1, std::min<int64_t>((M + 255) / 256, (148 * 8) / cb));
  const int64_t rows_per_block = (M + rb - 1) / rb;
  rb = (M + rows_per_block - 1) / rows_per_block;
  colsum_kernel<<<dim3(cb, (unsigned)rb), 256, 0, as_stream(stream)>>>(X, ld, M, ncols, rows_per_block, out);
  return check_launch("colsum_kernel");
}

// =====================================================================================================================
// SDF / albedo field, training path (neusky/fields/sdf_albedo_field.py:211-269 + nerfstudio SDFField.forward_geonetwork):
// the stages between the tf32 contractions, including the pieces of the normals' double backward.
// =====================================================================================================================

namespace nsk {
namespace train {

constexpr int H0_LD = 72;     // [x 3 | PE6 36 | hash 32 | pad 1]   (reference concatenation order, sdf_albedo_field.py / SDFField)
constexpr float TWO_PI_F = 6.283185307179586f;

// PE6(x): index d*6+k -> sin(2 pi 2^k x_d), 18 + d*6+k -> sin(2 pi 2^k x_d + pi/2)   [nerfstudio NeRFEncoding, SURVEY A.2]
__device__ __forceinline__ void pe6(const float x[3], float* out /*36*/) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float s = TWO_PI_F * x[d];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float a = s * (float)(1 << k);
      out[d * 6 + k] = sinf(a);
      out[18 + d * 6 + k] = sinf(a + 1.5707963267948966f);
    }
  }
}

// x [n,3] -> H0 [n,72] (x, PE, hash(pos), 0), tail [n, ld_tail] columns [0,39) = (x, PE) and column 39 = 0 (colour-net input
// tail; NULL = skip), pos [n,3] = (contract_inf(x) + 2) / 4, J [n,9] = d pos / d x
__global__ void __launch_bounds__(128)
sdf_inputs_fwd_kernel(const float* __restrict__ x, int64_t n, const float2* __restrict__ table, const float* __restrict__ scalings,
                      int L, int log2_T, float* __restrict__ H0, float* __restrict__ tail, int ld_tail, float* __restrict__ pos_out,
                      float* __restrict__ J_out) {
  const uint32_t mask = (1u << log2_T) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
    float pos[3], J[9], enc[36];
    sdf_contract(xv, pos, J);
    pe6(xv, enc);
    float* h = H0 + i * H0_LD;
    h[0] = xv[0]; h[1] = xv[1]; h[2] = xv[2];
#pragma unroll
    for (int c = 0; c < 36; ++c) h[3 + c] = enc[c];
    if (tail != nullptr) {
      float* t = tail + i * ld_tail;
      t[0] = xv[0]; t[1] = xv[1]; t[2] = xv[2];
#pragma unroll
      for (int c = 0; c < 36; ++c) t[3 + c] = enc[c];
      t[39] = 0.f;
    }
    for (int lev = 0; lev < L; ++lev) {
      const float s = scalings[lev];
      uint32_t idx[8];
      float ox, oy, oz;
      hash_corners(__fmul_rn(pos[0], s), __fmul_rn(pos[1], s), __fmul_rn(pos[2], s), mask, idx, ox, oy, oz);
      const float2* tl = table + ((size_t)lev << log2_T);
      float2 f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = __ldg(tl + idx[c]);
      const float2 v = hash_interp(f, ox, oy, oz);
      h[39 + 2 * lev] = v.x;
      h[40 + 2 * lev] = v.y;
    }
    h[71] = 0.f;
    pos_out[i * 3] = pos[0]; pos_out[i * 3 + 1] = pos[1]; pos_out[i * 3 + 2] = pos[2];
#pragma unroll
    for (int c = 0; c < 9; ++c) J_out[i * 9 + c] = J[c];
  }
}

// grad_x = G[:, 0:3] + PE'(x)^T G[:, 3:39] + J^T gpos      (the input stage of the reverse pass; G = d . / d H0)
__global__ void __launch_bounds__(128)
sdf_grad_assemble_kernel(const float* __restrict__ x, const float* __restrict__ G, const float* __restrict__ gpos,
                         const float* __restrict__ J, int64_t n, float* __restrict__ grad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* g = G + i * H0_LD;
    float out[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float s = TWO_PI_F * x[i * 3 + d];
      float acc = g[d];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float f = (float)(1 << k), a = s * f;
        // d sin(a)/dx = 2 pi f cos(a);  d sin(a + pi/2)/dx = 2 pi f cos(a + pi/2)
        acc += TWO_PI_F * f * (g[3 + d * 6 + k] * cosf(a) + g[3 + 18 + d * 6 + k] * cosf(a + 1.5707963267948966f));
      }
      out[d] = acc;
    }
    const float* Jr = J + i * 9;
    const float gp[3] = {gpos[i * 3], gpos[i * 3 + 1], gpos[i * 3 + 2]};
#pragma unroll
    for (int d = 0; d < 3; ++d) grad[i * 3 + d] = out[d] + Jr[0 * 3 + d] * gp[0] + Jr[1 * 3 + d] * gp[1] + Jr[2 * 3 + d] * gp[2];
  }
}

// transpose of the above for a cotangent c [n,3] on grad_x: dG[:, 0:3] = c, dG[:, 3:39] = PE'(x) c, dG[:, 71] = 0, cpos = J c
// (columns 39..70 of dG, the hash part, are filled from nsk_hash_encode_grad_x_bwd by the host)
__global__ void __launch_bounds__(128)
sdf_grad_assemble_bwd_kernel(const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ J, int64_t n,
                             float* __restrict__ dG, float* __restrict__ cpos) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float* g = dG + i * H0_LD;
    const float cv[3] = {c[i * 3], c[i * 3 + 1], c[i * 3 + 2]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float s = TWO_PI_F * x[i * 3 + d];
      g[d] = cv[d];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float f = (float)(1 << k), a = s * f;
        g[3 + d * 6 + k] = TWO_PI_F * f * cosf(a) * cv[d];
        g[3 + 18 + d * 6 + k] = TWO_PI_F * f * cosf(a + 1.5707963267948966f) * cv[d];
      }
    }
    g[71] = 0.f;
    const float* Jr = J + i * 9;
#pragma unroll
    for (int r = 0; r < 3; ++r) cpos[i * 3 + r] = Jr[r * 3 + 0] * cv[0] + Jr[r * 3 + 1] * cv[1] + Jr[r * 3 + 2] * cv[2];
  }
}

// softplus_100 derivatives through the OUTPUT a: s' = sigmoid(100 z) = 1 - exp(-100 a), s'' = 100 s' (1 - s')
__device__ __forceinline__ float sp_d1(float a) { return -expm1f(-100.0f * a); }
__device__ __forceinline__ float sp_d2(float a) { const float s = sp_d1(a); return 100.0f * s * (1.0f - s); }

// pointwise family over [n,256] fp32 tensors (a, b, c, d, out contiguous), w [256] column vector, s [n] row vector
//   0  out = w[col] * sp'(a)                          G2 of the reverse pass
//   1  out = a * sp'(b)
//   2  out = a * sp'(b) + c * d * sp''(b)             (c == NULL: second term dropped; a == NULL: first term dropped)
//   3  out = a * sp'(b) + c * w[col] * sp''(b)        (same NULL rules)
//   4  out = a + s[row] * w[col]                      (a == NULL: outer product only)
__global__ void __launch_bounds__(256)
ew256_kernel(int op, int64_t n, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
             const float* __restrict__ d, const float* __restrict__ w, const float* __restrict__ s, float* __restrict__ out) {
  const int64_t total = n * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 6;
    const int col = (int)(i & 63) * 4;
    const int64_t o = r * 256 + col;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 va = a ? *reinterpret_cast<const float4*>(a + o) : z4;
    const float4 vb = b ? *reinterpret_cast<const float4*>(b + o) : z4;
    const float4 vc = c ? *reinterpret_cast<const float4*>(c + o) : z4;
    const float4 vd = d ? *reinterpret_cast<const float4*>(d + o) : z4;
    const float4 vw = w ? *reinterpret_cast<const float4*>(w + col) : z4;
    const float sr = s ? s[r] : 0.f;
    float4 res;
#define NSK_EW(k)                                                                  \
  {                                                                                \
    float v;                                                                       \
    if (op == 0) v = vw.k * sp_d1(va.k);                                           \
    else if (op == 1) v = va.k * sp_d1(vb.k);                                      \
    else if (op == 2) v = va.k * sp_d1(vb.k) + (c ? vc.k * vd.k * sp_d2(vb.k) : 0.f); \
    else if (op == 3) v = va.k * sp_d1(vb.k) + (c ? vc.k * vw.k * sp_d2(vb.k) : 0.f); \
    else v = va.k + sr * vw.k;                                                     \
    res.k = v;                                                                     \
  }
    NSK_EW(x) NSK_EW(y) NSK_EW(z) NSK_EW(w)
#undef NSK_EW
    *reinterpret_cast<float4*>(out + o) = res;
  }
}

// out[r] = X[r, :256] . w + b   (exact fp32; the sdf head)
__global__ void __launch_bounds__(256)
rowdot256_kernel(const float* __restrict__ X, const float* __restrict__ w, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4 w0 = *reinterpret_cast<const float4*>(w + lane * 8), w1 = *reinterpret_cast<const float4*>(w + lane * 8 + 4);
  const float bias = b ? b[0] : 0.f;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4 x0 = *reinterpret_cast<const float4*>(X + r * 256 + lane * 8), x1 = *reinterpret_cast<const float4*>(X + r * 256 + lane * 8 + 4);
    float s = x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
    s = warp_sum(s);
    if (lane == 0) out[r] = s + bias;
  }
}

// out[c] += sum_r v[r] * X[r, c]
__global__ void __launch_bounds__(256)
colsum_w_kernel(const float* __restrict__ X, int ld, const float* __restrict__ v, int64_t M, int ncols, int64_t rows_per_block,
                float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (c >= ncols) return;
  float s0 = 0.f, s1 = 0.f;
  int64_t r = r0;
  for (; r + 1 < r1; r += 2) {
    s0 = fmaf(v[r], X[r * ld + c], s0);
    s1 = fmaf(v[r + 1], X[(r + 1) * ld + c], s1);
  }
  if (r < r1) s0 = fmaf(v[r], X[r * ld + c], s0);
  atomicAdd(out + c, s0 + s1);
}

}  // namespace train
}  // namespace nsk

extern "C" int nsk_sdf_inputs_fwd(const float* x, int64_t n, const float* table, const float* scalings, int num_levels,
                                  int log2_T, float* H0, float* tail, int ld_tail, float* pos, float* J, void* stream) {
  NSK_REQUIRE(x && table && scalings && H0 && pos && J, "nsk_sdf_inputs_fwd: null pointer");
  NSK_REQUIRE(num_levels == 16 && log2_T > 0 && log2_T < 31, "nsk_sdf_inputs_fwd: hash grid shape");
  NSK_REQUIRE(tail == nullptr || ld_tail >= 40, "nsk_sdf_inputs_fwd: ld_tail");
  if (n == 0) return 0;
  sdf_inputs_fwd_kernel<<<grid_for(n, 128, 148 * 16), 128, 0, as_stream(stream)>>>(x, n, reinterpret_cast<const float2*>(table), scalings,
                                                                                 num_levels, log2_T, H0, tail, ld_tail, pos, J);
  return check_launch("sdf_inputs_fwd_kernel");
}

extern "C" int nsk_sdf_grad_assemble(const float* x, const float* G, const float* gpos, const float* J, int64_t n, float* grad,
                                     void* stream) {
  NSK_REQUIRE(x && G && gpos && J && grad, "nsk_sdf_grad_assemble: null pointer");
  if (n == 0) return 0;
  sdf_grad_assemble_kernel<<<grid_for(n, 128, 148 * 16), 128, 0, as_stream(stream)>>>(x, G, gpos, J, n, grad);
  return check_launch("sdf_grad_assemble_kernel");
}

extern "C" int nsk_sdf_grad_assemble_bwd(const float* x, const float* c, const float* J, int64_t n, float* dG, float* cpos,
                                         void* stream) {
  NSK_REQUIRE(x && c && J && dG && cpos, "nsk_sdf_grad_assemble_bwd: null pointer");
  if (n == 0) return 0;
  sdf_grad_assemble_bwd_kernel<<<grid_for(n, 128, 148 * 16), 128, 0, as_stream(stream)>>>(x, c, J, n, dG, cpos);
  return check_launch("sdf_grad_assemble_bwd_kernel");
}

extern "C" int nsk_ew256(int op, int64_t n, const float* a, const float* b, const float* c, const float* d, const float* w,
                         const float* s, float* out, void* stream) {
  NSK_REQUIRE(op >= 0 && op <= 4 && out, "nsk_ew256: op / out");
  NSK_REQUIRE(op != 0 || (a && w), "nsk_ew256 op 0 needs a, w");
  NSK_REQUIRE(op != 1 || (a && b), "nsk_ew256 op 1 needs a, b");
  NSK_REQUIRE(op != 2 || (b && (!c || d)), "nsk_ew256 op 2 needs b (and d with c)");
  NSK_REQUIRE(op != 3 || (b && (!c || w)), "nsk_ew256 op 3 needs b (and w with c)");
  NSK_REQUIRE(op != 4 || (w && s), "nsk_ew256 op 4 needs w, s");
  if (n == 0) return 0;
  ew256_kernel<<<grid_for(n * 64, 256), 256, 0, as_stream(stream)>>>(op, n, a, b, c, d, w, s, out);
  return check_launch("ew256_kernel");
}

extern "C" int nsk_rowdot256(const float* X, const float* w, const float* b, int64_t n, float* out, void* stream) {
  NSK_REQUIRE(X && w && out, "nsk_rowdot256: null pointer");
  if (n == 0) return 0;
  rowdot256_kernel<<<grid_for(n * 32, 256), 256, 0, as_stream(stream)>>>(X, w, b, n, out);
  return check_launch("rowdot256_kernel");
}

extern "C" int nsk_colsum_w(const float* X, int ld, const float* v, int64_t M, int ncols, float* out, void* stream) {
  NSK_REQUIRE(X && v && out && ld >= ncols && ncols > 0, "nsk_colsum_w: arguments");
  if (M == 0) return 0;
  const int cb = (ncols + 255) / 256;
  int64_t rb = std::max<int64_t>(1, std::min<int64_t>((M + 255) / 256, (148 * 8) / cb));
  const int64_t rows_per_block = (M + rb - 1) / rb;
  rb = (M + rows_per_block - 1) / rows_per_block;
  colsum_w_kernel<<<dim3(cb, (unsigned)rb), 256, 0, as_stream(stream)>>>(X, ld, v, M, ncols, rows_per_block, out);
  return check_launch("colsum_w_kernel");
}

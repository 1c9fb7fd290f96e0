// Device helpers shared by the SDF / albedo field kernels (sdf_field_simt.cu, sdf_field_tc.cu).
#pragma once
#include "nsk_common.cuh"

namespace nsk {

constexpr int SDF_HID = 256;
constexpr int SDF_IN = 71;        // 3 + 36 + 32
constexpr int SDF_CIN = 295;      // 3 + 36 + 256
constexpr int SDF_LEVELS = 16;


// nerfstudio SceneContraction(order=inf) [NS-mem A.4] followed by (p + 2) / 4; J = d pos / d x (row-major 3x3).
__device__ __forceinline__ void sdf_contract(const float x[3], float pos[3], float J[9]) {
  const float ax = fabsf(x[0]), ay = fabsf(x[1]), az = fabsf(x[2]);
  const float mag = fmaxf(ax, fmaxf(ay, az));
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = 0.f;
  if (mag < 1.0f) {
    pos[0] = (x[0] + 2.0f) / 4.0f; pos[1] = (x[1] + 2.0f) / 4.0f; pos[2] = (x[2] + 2.0f) / 4.0f;
    J[0] = J[4] = J[8] = 0.25f;
    return;
  }
  const int im = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
  const float inv = 1.0f / mag;
  const float h = (2.0f - inv) * inv;                       // (2 - 1/mag) / mag
  const float dh = -2.0f * inv * inv + 2.0f * inv * inv * inv;
  const float sg = x[im] >= 0.f ? 1.0f : -1.0f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pos[j] = ((2.0f - inv) * (x[j] * inv) + 2.0f) / 4.0f;
    J[j * 3 + j] += 0.25f * h;
    J[j * 3 + im] += 0.25f * x[j] * dh * sg;
  }
}

// d interp / d (ox, oy, oz) for one feature channel, corner order of hash_corners()
__device__ __forceinline__ void hash_interp_grad(const float f[8], float ox, float oy, float oz, float d[3]) {
  const float f03 = f[0] * ox + f[3] * (1.f - ox), f12 = f[1] * ox + f[2] * (1.f - ox);
  const float f56 = f[5] * ox + f[6] * (1.f - ox), f47 = f[4] * ox + f[7] * (1.f - ox);
  const float f0312 = f03 * oy + f12 * (1.f - oy), f4756 = f47 * oy + f56 * (1.f - oy);
  d[2] = f0312 - f4756;
  d[1] = (f03 - f12) * oz + (f47 - f56) * (1.f - oz);
  d[0] = ((f[0] - f[3]) * oy + (f[1] - f[2]) * (1.f - oy)) * oz + ((f[4] - f[7]) * oy + (f[5] - f[6]) * (1.f - oy)) * (1.f - oz);
}

}  // namespace nsk

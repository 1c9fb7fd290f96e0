// Stand-alone probes of the Blackwell primitives the tensor-core shading kernel is built from.
// Each probe validates one mechanism against a CPU result and/or measures its throughput, so the
// kernel design rests on measured facts (B200, sm_100a) rather than assumptions.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_probe tools/tc_probe.cu
//   run  : ./tc_probe <probe> [args]     (see main())
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>

#include "../neusky_b200/csrc/tc_util.cuh"

using namespace nsk::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

// ------------------------------------------------------------------------------------------------
// probe 1/2: one tile D[128 x N] = A[128 x K] * B[N x K]^T, A from smem (SS) or TMEM (TS)
// ------------------------------------------------------------------------------------------------
template <bool A_IN_TMEM>
__global__ void __launch_bounds__(128) mma_tile_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                                       int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;                  // 128 x K fp16, canonical layout
  uint8_t* sB = smem + 128 * K * 2;    // N x K fp16
  const int t = threadIdx.x, warp = t >> 5;
  for (int e = t; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    *reinterpret_cast<__half*>(sA + tile_off(128, r, k)) = A[e];
  }
  for (int e = t; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    *reinterpret_cast<__half*>(sB + tile_off(N, r, k)) = B[e];
  }
  if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t acc = tmem;            // columns [0, N)
  const uint32_t a_tm = tmem + 256;     // columns [256, 256 + K/2): A as packed fp16 pairs
  if (A_IN_TMEM) {
    // thread t owns row t (TMEM lane t): column c holds elements (2c, 2c+1)
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const __half2 h = __halves2half2(A[t * K + 2 * (c0 + j)], A[t * K + 2 * (c0 + j) + 1]);
        v[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st8(a_tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (t == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint64_t bd = make_smem_desc(smem_u32(sB) + (k0 / 8) * N * 16, N * 16, 128);
      if (A_IN_TMEM) {
        umma_ts(acc, a_tm + k0 / 2, bd, idesc, k0 > 0);
      } else {
        const uint64_t ad = make_smem_desc(smem_u32(sA) + (k0 / 8) * 128 * 16, 128 * 16, 128);
        umma_ss(acc, ad, bd, idesc, k0 > 0);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)t * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static int probe_mma(bool ts, int N, int K) {
  std::vector<__half> hA(128 * K), hB((size_t)N * K);
  std::vector<float> fA(128 * K), fB((size_t)N * K), ref((size_t)128 * N), got((size_t)128 * N);
  srand(1234 + N + K);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 9) - 4) / 4.0f; hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 9) - 4) / 8.0f; hB[i] = __float2half(fB[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[(size_t)n * K + k];
      ref[(size_t)m * N + n] = s;
    }
  __half *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, got.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, got.size() * 4));
  const size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  if (ts) {
    CK(cudaFuncSetAttribute(mma_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_tile_kernel<true><<<1, 128, smem>>>(dA, dB, dD, N, K);
  } else {
    CK(cudaFuncSetAttribute(mma_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_tile_kernel<false><<<1, 128, smem>>>(dA, dB, dD, N, K);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    const double e = fabs((double)got[i] - ref[i]);
    if (!(e <= 1e-5)) { if (bad < 5) printf("  mismatch at m=%zu n=%zu got %f ref %f\n", i / N, i % N, got[i], ref[i]); ++bad; }
    if (e > maxerr) maxerr = e;
  }
  printf("probe mma_%s N=%d K=%d : %s (max err %.3g, %d bad)\n", ts ? "ts" : "ss", N, K, bad ? "FAIL" : "PASS", maxerr, bad);
  return bad != 0;
}

// ------------------------------------------------------------------------------------------------
// probe 3: MMA issue throughput on every SM (cycles per 128xNx16 MMA), SS vs TS, with or without
// concurrent shared-memory traffic from 4 "epilogue" warps.
// ------------------------------------------------------------------------------------------------
__device__ int g_commit_every = 0;   // power of two; 0 = a single commit at the end
template <bool A_IN_TMEM>
__global__ void __launch_bounds__(256) mma_rate_kernel(int N, int iters, int smem_noise, long long* __restrict__ cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  // operands: A 128x256, B 256x256 (zeros are fine for timing)
  for (int e = t; e < (128 + 256) * 256 * 2 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 128 * 256 * 2;
  if (t == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k0 = (i & 15) * 16;
      const uint64_t bd = make_smem_desc(sB + (k0 / 8) * N * 16, N * 16, 128);
      if (A_IN_TMEM) umma_ts(tmem, tmem + 256 + k0 / 2, bd, idesc, 1);
      else umma_ss(tmem, make_smem_desc(sA + (k0 / 8) * 128 * 16, 128 * 16, 128), bd, idesc, 1);
      if (g_commit_every > 0 && ((i + 1) & (g_commit_every - 1)) == 0) umma_commit(smem_u32(&bar2));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (warp >= 4 && smem_noise) {
    // emulate epilogue traffic: 16-byte stores + loads on a scratch region while the MMAs run
    uint8_t* scratch = smem + (128 + 256) * 256 * 2;
    float acc = 0.f;
    const int lt = t - 128;
    for (int i = 0; i < iters * smem_noise; ++i) {
      uint4* p = reinterpret_cast<uint4*>(scratch) + ((i * 128 + lt) & 1023);
      *p = make_uint4(i, t, 0, 0);
      acc += (float)p->x;
    }
    if (acc == 12345.f) *sink = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static void probe_rate(bool ts, int N, int noise, int commit_every = 0) {
  const int iters = 4096, grid = 148;
  CK(cudaMemcpyToSymbol(g_commit_every, &commit_every, sizeof(int)));
  long long* d; float* sink;
  CK(cudaMalloc(&d, grid * sizeof(long long))); CK(cudaMalloc(&sink, 4));
  const size_t smem = (size_t)(128 + 256) * 256 * 2 + 16384 + 1024;
  auto k = ts ? mma_rate_kernel<true> : mma_rate_kernel<false>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 256, smem>>>(N, iters, noise, d, sink);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, 256, smem>>>(N, iters, noise, d, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= grid;
  const double flops = 2.0 * 128 * N * 16 * (double)iters * grid;
  printf("probe rate_%s N=%d noise=%d commit_every=%d : %.1f cycles/MMA (ideal %d), kernel %.3f ms -> %.0f TFLOP/s\n", ts ? "ts" : "ss", N, noise, commit_every,
         mean / iters, N / 2, ms, flops / (ms * 1e-3) / 1e12);
}

// ------------------------------------------------------------------------------------------------
// probe 4: bulk-copy (TMA engine) streaming of an L2-resident weight blob through a smem ring:
// every CTA streams the same `blob_bytes` region `passes` times -- the access pattern of the
// per-tile weight stream.  Reports aggregate L2->SM bandwidth.
// ------------------------------------------------------------------------------------------------
constexpr int RING = 3;
__global__ void __launch_bounds__(128) stream_kernel(const uint8_t* __restrict__ blob, int stages_per_pass, int passes, int stage_bytes,
                                                     unsigned* __restrict__ check) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[RING], empty[RING];
  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int total = stages_per_pass * passes;
  if (t == 0) {  // producer
    const uint64_t pol = l2_policy_evict_last();
    for (int i = 0; i < total; ++i) {
      const int s = i % RING, ph = (i / RING) & 1;
      mbar_wait(smem_u32(&empty[s]), ph ^ 1);
      mbar_arrive_expect_tx(smem_u32(&full[s]), stage_bytes);
      bulk_g2s_hint(smem_u32(smem + (size_t)s * stage_bytes), blob + (size_t)(i % stages_per_pass) * stage_bytes, stage_bytes,
                    smem_u32(&full[s]), pol);
    }
  } else if (t == 32) {  // consumer: touch one word per stage, release
    unsigned acc = 0;
    for (int i = 0; i < total; ++i) {
      const int s = i % RING, ph = (i / RING) & 1;
      mbar_wait(smem_u32(&full[s]), ph);
      acc += *reinterpret_cast<volatile unsigned*>(smem + (size_t)s * stage_bytes + 4 * (i & 63));
      mbar_arrive(smem_u32(&empty[s]));
    }
    check[blockIdx.x] = acc;
  }
}

static int probe_stream(int stage_kb, int grid) {
  const int stage_bytes = stage_kb * 1024;
  const int blob_bytes = 2400 * 1024 / stage_bytes * stage_bytes;
  const int stages_per_pass = blob_bytes / stage_bytes, passes = 40;
  std::vector<unsigned> h(blob_bytes / 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned)i * 2654435761u;
  uint8_t* d; unsigned* chk;
  CK(cudaMalloc(&d, blob_bytes)); CK(cudaMalloc(&chk, grid * 4));
  CK(cudaMemcpy(d, h.data(), blob_bytes, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)RING * stage_bytes + 1024;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stream_kernel<<<grid, 128, smem>>>(d, stages_per_pass, 2, stage_bytes, chk);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  stream_kernel<<<grid, 128, smem>>>(d, stages_per_pass, passes, stage_bytes, chk);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned> got(grid);
  CK(cudaMemcpy(got.data(), chk, grid * 4, cudaMemcpyDeviceToHost));
  unsigned ref = 0;
  for (int i = 0; i < stages_per_pass * passes; ++i) ref += h[((size_t)(i % stages_per_pass) * stage_bytes + 4 * (i & 63)) / 4];
  int bad = 0; for (auto v : got) bad += (v != ref);
  const double bytes = (double)blob_bytes * passes * grid;
  printf("probe stream stage=%dKB grid=%d : %s, %.3f ms, %.2f TB/s aggregate (%.1f GB/s per CTA)\n", stage_kb, grid, bad ? "FAIL" : "PASS", ms,
         bytes / (ms * 1e-3) / 1e12, bytes / grid / (ms * 1e-3) / 1e9);
  return bad != 0;
}

// ------------------------------------------------------------------------------------------------
// probe 5: TMEM load throughput (4 warps x 32x32b.x32) and a sin/FMA epilogue cost estimate
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ldtm_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    tmem_ld32(tmem + (i & 7) * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __sinf(__uint_as_float(v[j]) * 0.001f + 1.0f);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 1234.5f) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

static void probe_ldtm() {
  long long* d; float* sink; CK(cudaMalloc(&d, 148 * 8)); CK(cudaMalloc(&sink, 4));
  const int iters = 2000;
  ldtm_kernel<<<148, 128>>>(iters, d, sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= 148;
  printf("probe ldtm : %.1f cycles per (4 warps x [32 lanes x 32 cols] load + 32 sin/FMA per thread) -> %.2f cycles per 128x32 element block\n",
         mean / iters, mean / iters);
}

// ------------------------------------------------------------------------------------------------
// probe 6: issue rate of tcgen05.mma.cta_group::2 (M = 256 across a CTA pair), A and B from shared memory.
// commit_every > 0: a multicast tcgen05.commit after every `commit_every` MMAs (the per-stage pattern of the kernel).
// ------------------------------------------------------------------------------------------------
// local-only commit of a cta_group::2 MMA group (no multicast): arrives on the mbarrier of the executing CTA only
__device__ __forceinline__ void umma_commit2_local(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_rate2_kernel(int N, int iters, int commit_every, int mode, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int e = t; e < (128 + 128) * 256 * 2 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc2<512>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int Nh = N / 2;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 128 * 256 * 2;
  uint32_t peer_bar2;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar2) : "r"(smem_u32(&bar2)), "r"(1));
  if (t == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, N);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k0 = (i & 15) * 16;
      if (mode == 3) {
        // A operand from TMEM (columns [256, 384): K = 256 halves of each CTA's 128 rows), B from shared memory: the ".ts" form of the pair MMA
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
            "r"(tmem + 256 + k0 / 2), "l"(make_smem_desc(sB + (k0 / 8) * Nh * 16, Nh * 16, 128)), "r"(idesc), "r"(1)
            : "memory");
      } else
      umma_ss2(tmem, make_smem_desc(sA + (k0 / 8) * 128 * 16, 128 * 16, 128), make_smem_desc(sB + (k0 / 8) * Nh * 16, Nh * 16, 128), idesc, 1);
      if (commit_every > 0 && ((i + 1) & (commit_every - 1)) == 0) {
        if (mode == 0) umma_commit2(smem_u32(&bar2));                 // multicast to both CTAs
        else if (mode == 1) umma_commit2_local(smem_u32(&bar2));      // leader's barrier only
        else { umma_commit2_local(smem_u32(&bar2)); umma_commit2_local(peer_bar2); }   // two unicast commits: local + peer address
      }
    }
    umma_commit2(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x >> 1] = clock64() - t0;
  } else if (t == 0) {
    mbar_wait(smem_u32(&bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2<512>(tmem);
}

static void probe_rate2(int N, int commit_every, int mode) {
  const int iters = 4096, grid = 148;
  long long* d;
  CK(cudaMalloc(&d, grid / 2 * sizeof(long long)));
  const size_t smem = (size_t)(128 + 128) * 256 * 2 + 1024;
  CK(cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) mma_rate2_kernel<<<grid, 128, smem>>>(N, iters, commit_every, mode, d);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mma_rate2_kernel<<<grid, 128, smem>>>(N, iters, commit_every, mode, d);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid / 2);
  CK(cudaMemcpy(h.data(), d, grid / 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= (grid / 2);
  const double flops = 2.0 * 256 * N * 16 * (double)iters * (grid / 2);
  printf("probe rate2 (cta_group::2, M=256) N=%d commit_every=%d mode=%d : %.1f cycles/MMA (ideal %d), kernel %.3f ms -> %.0f TFLOP/s\n", N, commit_every, mode,
         mean / iters, N / 2, ms, flops / (ms * 1e-3) / 1e12);
}


// ------------------------------------------------------------------------------------------------
// probe 7: pure TMEM read throughput / latency: `warps` warps (warp w reads lane quadrant w % 4), 32x32b.xN loads, `depth` loads
// in flight per warp before a tcgen05.wait::ld; no math beyond an xor fold.  Reports bytes per clock per SM and cycles per load group.
// ------------------------------------------------------------------------------------------------
template <int X>
__global__ void __launch_bounds__(512) ldtm2_kernel(int iters, int depth, long long* cycles, unsigned* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[4][X];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (d < depth) {
        if (X == 8) tmem_ld8(tmem + ((i + d) & 1) * 32 + d * 8, reinterpret_cast<uint32_t(&)[8]>(v[d]));
        if (X == 16) tmem_ld16(tmem + ((i + d) & 1) * 32 + (d & 1) * 16, reinterpret_cast<uint32_t(&)[16]>(v[d]));
        if (X == 32) tmem_ld32(tmem + ((i + d) & 1) * 32, reinterpret_cast<uint32_t(&)[32]>(v[d]));
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < depth) {
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= v[d][j];
      }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

static void probe_ldtm2(int warps, int x, int depth) {
  long long* d; unsigned* sink; CK(cudaMalloc(&d, 148 * 8)); CK(cudaMalloc(&sink, 4));
  const int iters = 4000;
  if (x == 8) ldtm2_kernel<8><<<148, warps * 32>>>(iters, depth, d, sink);
  else if (x == 16) ldtm2_kernel<16><<<148, warps * 32>>>(iters, depth, d, sink);
  else ldtm2_kernel<32><<<148, warps * 32>>>(iters, depth, d, sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= 148;
  const double bytes = (double)iters * depth * x * 4 * 32 * warps;
  printf("probe ldtm2 warps=%d x%d depth=%d : %.1f cycles per group, %.1f B/clk per SM\n", warps, x, depth, mean / iters, bytes / mean);
}


// ------------------------------------------------------------------------------------------------
// probe 8: TMEM read latency / throughput UNDER a running MMA stream: thread 0 issues tcgen05.mma (M = 128, N, K = 16, accumulator in
// columns [0, N)) back to back while warps 4.. read columns [256, 512) with 32x32b.x8 loads, `depth` loads in flight.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(640) ldtm3_kernel(int N, int mma_iters, int ld_iters, int depth, long long* cycles, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int stop;
  const int t = threadIdx.x, warp = t >> 5;
  for (int e = t; e < (128 + 256) * 256 * 2 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 128 * 256 * 2;
  if (t == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int i = 0; i < mma_iters; ++i) {
      const int k0 = (i & 15) * 16;
      umma_ss(tmem, make_smem_desc(sA + (k0 / 8) * 128 * 16, 128 * 16, 128), make_smem_desc(sB + (k0 / 8) * N * 16, N * 16, 128), idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  } else if (warp >= 4) {
    const int w = warp - 4;
    const uint32_t ta = tmem + ((uint32_t)((w & 3) * 32) << 16) + 256 + (uint32_t)((w >> 2) * 64);
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      uint32_t v[4][8];
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (d < depth) tmem_ld8(ta + ((i + d) & 1) * 32 + d * 8, v[d]);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (d < depth) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc ^= v[d][j];
        }
    }
    const long long t1 = clock64();
    if (t == 128) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static void probe_ldtm3(int N, int warps, int depth) {
  long long* d; unsigned* sink; CK(cudaMalloc(&d, 148 * 8)); CK(cudaMalloc(&sink, 4));
  const int ld_iters = 4000;
  const size_t smem = (size_t)(128 + 256) * 256 * 2 + 1024;
  CK(cudaFuncSetAttribute(ldtm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // enough MMAs to outlast the loads: ~ld_iters * 400 cycles / (N / 2 cycles per MMA)
  const int mma_iters = N > 0 ? (int)(ld_iters * 600.0 / (N / 2)) : 0;
  ldtm3_kernel<<<148, (4 + warps) * 32, smem>>>(N > 0 ? N : 256, mma_iters, ld_iters, depth, d, sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= 148;
  printf("probe ldtm3 mma N=%d warps=%d depth=%d (x8) : %.1f cycles per group, %.1f B/clk per SM\n", N, warps, depth, mean / ld_iters,
         (double)ld_iters * depth * 8 * 4 * 32 * warps / mean);
}


// ------------------------------------------------------------------------------------------------
// probe 9: throughput of the sine paths available to an epilogue: __sinf (FMUL + MUFU.SIN), a degree-11 odd polynomial after an exact
// period reduction (FMA pipe only), and a 50/50 mix, with `warps` warps per SM (warps / 4 per scheduler), 32 independent values per thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sin_poly(float arg) {
  // sin(arg) = sin(2 pi r), r = arg / 2 pi - rint(arg / 2 pi) in [-0.5, 0.5]; odd minimax-style polynomial in r (Taylor coefficients of sin(2 pi r)
  // re-fitted would do better; Taylor to r^13 is ~2e-7 absolute on the interval after folding to [-0.25, 0.25])
  float u = arg * 0.15915494309189535f;
  float r = u - rintf(u);
  // fold to [-0.25, 0.25]: sin(2 pi r) = sin(2 pi (0.5 sgn(r) - r))
  float rf = copysignf(0.5f, r) - r;
  r = fabsf(r) > 0.25f ? rf : r;
  const float x = 6.283185307179586f * r, x2 = x * x;
  float p = -2.5052108e-8f;
  p = fmaf(p, x2, 2.7557319e-6f);
  p = fmaf(p, x2, -1.9841270e-4f);
  p = fmaf(p, x2, 8.3333333e-3f);
  p = fmaf(p, x2, -1.6666667e-1f);
  return fmaf(p * x2, x, x);
}
template <int MODE>
__global__ void __launch_bounds__(512) sin_rate_kernel(int iters, long long* cycles, float* sink, float* err) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.37f * (threadIdx.x + 1) + 1.7f * j;
  float acc = 0.f, e = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = v[j] + (float)i * 0.01f;
      float s;
      if (MODE == 0) s = __sinf(a);
      else if (MODE == 1) s = sin_poly(a);
      else s = (j & 1) ? __sinf(a) : sin_poly(a);
      acc += s;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  // accuracy of the polynomial path against sinf on a sweep of arguments in the range the FiLM layers see (|arg| < 64)
  for (int k = 0; k < 64; ++k) {
    const float a = -64.f + (threadIdx.x * 64 + k) * (128.f / (512 * 64));
    e = fmaxf(e, fabsf(sin_poly(a) - (float)sin((double)a)));
  }
  if (acc == 1234.5f) *sink = acc;
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(e));
}
static void probe_sin(int mode, int warps) {
  long long* d; float* sink; float* err; CK(cudaMalloc(&d, 148 * 8)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  const int iters = 2000;
  if (mode == 0) sin_rate_kernel<0><<<148, warps * 32>>>(iters, d, sink, err);
  else if (mode == 1) sin_rate_kernel<1><<<148, warps * 32>>>(iters, d, sink, err);
  else sin_rate_kernel<2><<<148, warps * 32>>>(iters, d, sink, err);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148); float herr = 0;
  CK(cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto v : h) mean += (double)v; mean /= 148;
  printf("probe sin mode=%d (0 __sinf, 1 polynomial, 2 half/half) warps=%d : %.2f cycles per warp-sine per scheduler (%.1f sines/clk/SM), polynomial max |err| %.2e\n", mode, warps,
         mean / iters / 32 / (warps / 4.0), 32.0 * warps * 32 * iters / mean, herr);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  int fails = 0;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, cc %d.%d, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor, prop.sharedMemPerBlockOptin);
  if (!strcmp(what, "mma_ss")) { fails += probe_mma(false, atoi(argv[2]), atoi(argv[3])); }
  else if (!strcmp(what, "mma_ts")) { fails += probe_mma(true, atoi(argv[2]), atoi(argv[3])); }
  else if (!strcmp(what, "rate")) { probe_rate(atoi(argv[2]) != 0, atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 0); }
  else if (!strcmp(what, "stream")) { fails += probe_stream(atoi(argv[2]), atoi(argv[3])); }
  else if (!strcmp(what, "ldtm")) { probe_ldtm(); }
  else if (!strcmp(what, "sin")) { probe_sin(atoi(argv[2]), atoi(argv[3])); }
  else if (!strcmp(what, "ldtm3")) { probe_ldtm3(atoi(argv[2]), atoi(argv[3]), atoi(argv[4])); }
  else if (!strcmp(what, "ldtm2")) { probe_ldtm2(atoi(argv[2]), atoi(argv[3]), atoi(argv[4])); }
  else if (!strcmp(what, "rate2")) { probe_rate2(atoi(argv[2]), atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 0); }
  else { printf("unknown probe %s\n", what); return 2; }
  return fails ? 1 : 0;
}

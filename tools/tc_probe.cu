// tcgen05 bring-up probe: one 128xN += 128xK * (NxK)^T bf16 MMA chain per launch, checked against a
// host reference.  Each invocation tests ONE variant (so a faulting variant cannot poison the rest):
//   tc_probe <mode> <N> <K> <swap>
//     mode 0: A and B from shared memory (SS), canonical K-major no-swizzle layout
//     mode 1: A from tensor memory (TS), lane = row, two bf16 per 32-bit column
//     mode 2: SS, bf16 hi/lo split on both operands (3 MMAs per K step) vs exact fp32 product
//     mode 3: "dX" form: D[128,K] = A[128,N] * B[N,K] with B read MN-major from the SAME canonical
//             [N rows][K cols] tile a forward GEMM uses (no transposed copy)
//     mode 4: "dW" form: D[64,K] = A[128,64]^T * X[128,K], both operands MN-major views of canonical
//             row-per-point tiles (reduction over the 128 points), M=64 accumulator layout
//     mode 5: SS, A = bf16, B = fp16 in the same MMA (mixed operand formats)
//     mode 6: SS, fp16 hi/lo split on both operands (3 MMAs per K step) vs exact fp32 product
//     swap 1: exchange the LBO / SBO fields of the descriptors (layout-convention check)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../fruitnerf_b200/csrc/fnr_tcgen05.cuh"

using namespace fnr::tc;

__device__ int g_timeout;

__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 28)) {
      g_timeout = 1;
      return false;
    }
  }
  return true;
}

// A_hi/A_lo: [128][K] fp32 source; B: [N][K] fp32 source; D: [128][N]
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    int N, int K, int mode, int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const int chunks = K / 8;
  uint8_t* sA_hi = smem;                                  // 128 x K bf16
  uint8_t* sA_lo = sA_hi + 128 * K * 2;
  uint8_t* sB_hi = sA_lo + 128 * K * 2;                   // N x K bf16
  uint8_t* sB_lo = sB_hi + N * K * 2;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (t == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // stage A (thread t = row t) and B (rows strided over threads) in the canonical layout
  const bool a_f16 = mode == 6, b_f16 = mode == 5 || mode == 6;
  for (int j = 0; j < chunks; ++j) {
    uint32_t hi[4], lo[4];
    for (int q = 0; q < 4; ++q) {
      float a0 = A[t * K + j * 8 + 2 * q], a1 = A[t * K + j * 8 + 2 * q + 1], h0, l0, h1, l1;
      if (a_f16) {
        split_f16(a0, h0, l0);
        split_f16(a1, h1, l1);
        hi[q] = pack_f16x2(h0, h1);
        lo[q] = pack_f16x2(l0, l1);
      } else {
        split_bf16(a0, h0, l0);
        split_bf16(a1, h1, l1);
        hi[q] = pack_bf16x2(h0, h1);
        lo[q] = pack_bf16x2(l0, l1);
      }
    }
    *reinterpret_cast<uint4*>(sA_hi + j * 128 * 16 + t * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sA_lo + j * 128 * 16 + t * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int n = t; n < N; n += 128)
    for (int j = 0; j < chunks; ++j) {
      uint32_t hi[4], lo[4];
      for (int q = 0; q < 4; ++q) {
        float b0 = B[n * K + j * 8 + 2 * q], b1 = B[n * K + j * 8 + 2 * q + 1], h0, l0, h1, l1;
        if (b_f16) {
          split_f16(b0, h0, l0);
          split_f16(b1, h1, l1);
          hi[q] = pack_f16x2(h0, h1);
          lo[q] = pack_f16x2(l0, l1);
        } else {
          split_bf16(b0, h0, l0);
          split_bf16(b1, h1, l1);
          hi[q] = pack_bf16x2(h0, h1);
          lo[q] = pack_bf16x2(l0, l1);
        }
      }
      *reinterpret_cast<uint4*>(sB_hi + j * N * 16 + n * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sB_lo + j * N * 16 + n * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t d_tmem = tmem;            // columns [0, N)
  const uint32_t a_tmem = tmem + 256;      // columns [256, 256 + K/2)
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

  if (mode == 1) {  // A (hi part) into tensor memory: row = lane, 2 bf16 per column
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int q = 0; q < 8; ++q) {
        float a0 = A[t * K + 2 * (c0 + q)], a1 = A[t * K + 2 * (c0 + q) + 1], h0, l0, h1, l1;
        split_bf16(a0, h0, l0);
        split_bf16(a1, h1, l1);
        v[q] = pack_bf16x2(h0, h1);
      }
      tmem_st8(a_tmem + lane_addr + c0, v);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }

  if (t == 0) {
    const uint32_t idesc = idesc_f32acc(128, N, a_f16 ? 0 : 1, b_f16 ? 0 : 1);
    const uint32_t a_lbo = swap ? 128 : 128 * 16, a_sbo = swap ? 128 * 16 : 128;
    const uint32_t b_lbo = swap ? 128 : N * 16, b_sbo = swap ? N * 16 : 128;
    bool acc = false;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ah = smem_desc(smem_u32(sA_hi) + ks * 2 * 128 * 16, a_lbo, a_sbo);
      const uint64_t al = smem_desc(smem_u32(sA_lo) + ks * 2 * 128 * 16, a_lbo, a_sbo);
      const uint64_t bh = smem_desc(smem_u32(sB_hi) + ks * 2 * N * 16, b_lbo, b_sbo);
      const uint64_t bl = smem_desc(smem_u32(sB_lo) + ks * 2 * N * 16, b_lbo, b_sbo);
      if (mode == 1) {
        mma_ts(d_tmem, a_tmem + ks * 8, bh, idesc, acc);
      } else {
        mma_ss(d_tmem, ah, bh, idesc, acc);
        if (mode == 2 || mode == 6) {
          mma_ss(d_tmem, al, bh, idesc, true);
          mma_ss(d_tmem, ah, bl, idesc, true);
        }
      }
      acc = true;
    }
    mma_commit(&bar);
  }
  const bool ok = wait_bounded(&bar, 0);
  fence_after_sync();
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(d_tmem + lane_addr + c0, v);
      tmem_ld_wait();
      for (int q = 0; q < 8; ++q) D[t * N + c0 + q] = __uint_as_float(v[q]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  r &= 0xFFFF0000u;
  float y;
  memcpy(&y, &r, 4);
  return y;
}

#include <cuda_fp16.h>
static float f16_round_host(float x) { return __half2float(__float2half_rn(x)); }

__host__ __device__ constexpr uint32_t idesc_major(int M, int N, int a_mn, int b_mn) {
  return idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// mode 3/4 kernel.  A: [128][KA] fp32 (row per point), B: [ROWSB][KB] fp32.
//   mode 3: D[128][KB] = A[128][ROWSB(=KA)] * B[ROWSB][KB]
//   mode 4: D[KA(=64)][KB] = A^T * B with ROWSB = 128
__global__ void __launch_bounds__(128) probe_mn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                       int KA, int ROWSB, int KB, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  uint8_t* sA = smem;                    // 128 x KA bf16 canonical (rows = points)
  uint8_t* sB = sA + 128 * KA * 2;       // ROWSB x KB bf16 canonical
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (t == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int j = 0; j < KA / 8; ++j) {
    uint32_t hi[4];
    for (int q = 0; q < 4; ++q) hi[q] = pack_bf16x2(A[t * KA + j * 8 + 2 * q], A[t * KA + j * 8 + 2 * q + 1]);
    *reinterpret_cast<uint4*>(sA + j * 128 * 16 + t * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  }
  for (int n = t; n < ROWSB; n += 128)
    for (int j = 0; j < KB / 8; ++j) {
      uint32_t hi[4];
      for (int q = 0; q < 4; ++q) hi[q] = pack_bf16x2(B[n * KB + j * 8 + 2 * q], B[n * KB + j * 8 + 2 * q + 1]);
      *reinterpret_cast<uint4*>(sB + j * ROWSB * 16 + n * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  if (t == 0) {
    if (mode == 3) {
      // A K-major: K' = KA (=ROWSB); B MN-major view of sB: N' = KB, K' = ROWSB
      const uint32_t idesc = idesc_major(128, KB, 0, 1);
      for (int ks = 0; ks < ROWSB / 16; ++ks) {
        const uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
        // MN-major: SBO = stride between 8-element MN chunks (= ROWSB*16), LBO = stride between 8-row K groups (=128)
        const uint64_t bd = smem_desc(smem_u32(sB) + ks * 256, 128, ROWSB * 16);
        mma_ss(tmem, ad, bd, idesc, ks > 0);
      }
    } else {
      // both MN-major, reduction over the 128 rows; M' = KA (64), N' = KB
      const uint32_t idesc = idesc_major(KA, KB, 1, 1);
      for (int ks = 0; ks < 128 / 16; ++ks) {
        const uint64_t ad = smem_desc(smem_u32(sA) + ks * 256, 128, 128 * 16);
        const uint64_t bd = smem_desc(smem_u32(sB) + ks * 256, 128, 128 * 16);
        mma_ss(tmem, ad, bd, idesc, ks > 0);
      }
    }
    mma_commit(&bar);
  }
  const bool ok = wait_bounded(&bar, 0);
  fence_after_sync();
  if (ok) {
    // dump all 128 lanes x KB columns (mode 4 uses the M=64 lane layout; the host decodes it)
    for (int c0 = 0; c0 < KB; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(tmem + lane_addr + c0, v);
      tmem_ld_wait();
      for (int q = 0; q < 8; ++q) D[t * KB + c0 + q] = __uint_as_float(v[q]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_mn(int mode, int KA, int KB) {
  const int ROWSB = mode == 3 ? KA : 128;
  std::vector<float> A(128 * KA), B(ROWSB * KB), D(128 * KB, -777.f);
  uint32_t s = 999u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
  };
  auto bf = [](float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    uint32_t r = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    float y;
    memcpy(&y, &r, 4);
    return y;
  };
  for (auto& v : A) v = bf(rnd() * 1.7f);
  for (auto& v : B) v = bf(rnd() * 0.9f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 128 * KA * 2 + (size_t)ROWSB * KB * 2;
  cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mn_kernel<<<1, 128, smem>>>(dA, dB, dD, KA, ROWSB, KB, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PROBE mode=%d KA=%d KB=%d : CUDA ERROR %s\n", mode, KA, KB, cudaGetErrorString(e));
    return 2;
  }
  int timeout = 0;
  cudaMemcpyFromSymbol(&timeout, g_timeout, sizeof(int));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  const int Mrows = mode == 3 ? 128 : KA;
  for (int m = 0; m < Mrows; ++m)
    for (int n = 0; n < KB; ++n) {
      double ref = 0;
      if (mode == 3)
        for (int k = 0; k < KA; ++k) ref += (double)A[m * KA + k] * B[k * KB + n];
      else
        for (int p = 0; p < 128; ++p) ref += (double)A[p * KA + m] * B[p * KB + n];
      // M=64 accumulators: row i lives in TMEM lane (i % 16) + 32 * (i / 16)
      const int lane = (mode == 3 || Mrows == 128) ? m : (m % 16) + 32 * (m / 16);
      max_err = fmax(max_err, fabs(ref - D[lane * KB + n]));
      max_ref = fmax(max_ref, fabs(ref));
    }
  printf("PROBE mode=%d KA=%d KB=%d : timeout=%d max_err=%.3e max_ref=%.3e rel=%.3e %s\n", mode, KA, KB, timeout, max_err, max_ref,
         max_err / max_ref, (!timeout && max_err / max_ref < 2e-5) ? "PASS" : "FAIL");
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && (atoi(argv[1]) == 3 || atoi(argv[1]) == 4))
    return run_mn(atoi(argv[1]), argc > 2 ? atoi(argv[2]) : 64, argc > 3 ? atoi(argv[3]) : 64);
  const int mode = argc > 1 ? atoi(argv[1]) : 0, N = argc > 2 ? atoi(argv[2]) : 64, K = argc > 3 ? atoi(argv[3]) : 32,
            swap = argc > 4 ? atoi(argv[4]) : 0;
  std::vector<float> A(128 * K), B(N * K), D(128 * N, -777.f);
  uint32_t s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
  };
  for (auto& v : A) v = rnd() * 1.7f;
  for (auto& v : B) v = rnd() * 0.9f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * 128 * K * 2 + 2 * (size_t)N * K * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, swap);
  cudaError_t e = cudaDeviceSynchronize();
  int timeout = 0;
  if (e == cudaSuccess) cudaMemcpyFromSymbol(&timeout, g_timeout, sizeof(int));
  if (e != cudaSuccess) {
    printf("PROBE mode=%d N=%d K=%d swap=%d : CUDA ERROR %s\n", mode, N, K, swap, cudaGetErrorString(e));
    return 2;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        const double a = (mode == 2 || mode == 6) ? A[m * K + k] : bf16_round(A[m * K + k]);
        const double b = (mode == 2 || mode == 6) ? B[n * K + k] : (mode == 5 ? f16_round_host(B[n * K + k]) : bf16_round(B[n * K + k]));
        ref += a * b;
      }
      max_err = fmax(max_err, fabs(ref - D[m * N + n]));
      max_ref = fmax(max_ref, fabs(ref));
    }
  const double tol = mode == 2 ? 2e-4 : (mode == 6 ? 2e-6 : 2e-5);
  printf("PROBE mode=%d N=%d K=%d swap=%d : timeout=%d max_err=%.3e max_ref=%.3e rel=%.3e %s\n", mode, N, K, swap, timeout, max_err,
         max_ref, max_err / max_ref, (!timeout && max_err / max_ref < tol) ? "PASS" : "FAIL");
  return 0;
}

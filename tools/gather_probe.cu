// Scattered 16-byte gather throughput per SM on sm_100a: which datapath is fastest for hash-table rows?
//   mode 0: ld.global.nc.v4 (LSU / L1TEX), registers
//   mode 1: cp.async.bulk 16 B global -> shared (TMA unit, bypasses L1TEX), completion on an mbarrier
//   mode 2: half of the rows through each path concurrently
//   mode 3: cp.async (LDGSTS) 16 B global -> shared
// gather_probe <mode> <log2_rows> <threads> <copies_per_thread_per_iter> <iters>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/gather_probe tools/gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../fruitnerf_b200/csrc/fnr_tcgen05.cuh"

using namespace fnr::tc;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ void bulk16(uint32_t dst, const void* src, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst), "l"(src), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldgsts16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(1024) probe(const float4* __restrict__ table, uint32_t mask, int per_thread, int iters, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t sbase = smem_u32(smem), sbar = smem_u32(&bar);
  float acc = 0.f;
  uint32_t phase = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t seed = (blockIdx.x * 7919u + it) * 2654435761u;
    const int n_bulk = MODE == 1 ? per_thread : (MODE == 2 ? per_thread / 2 : 0);
    if (n_bulk && tid == 0) expect_tx(sbar, (uint32_t)n_bulk * nt * 16u);
    if (n_bulk) __syncthreads();  // expect_tx before any complete_tx
    for (int j = 0; j < n_bulk; ++j) {
      const uint32_t r = mix(seed + tid * 64u + j) & mask;
      bulk16(sbase + (uint32_t)(j * nt + tid) * 16u, table + r, sbar);
    }
    if (MODE == 3) {
      for (int j = 0; j < per_thread; ++j) {
        const uint32_t r = mix(seed + tid * 64u + j) & mask;
        ldgsts16(sbase + (uint32_t)(j * nt + tid) * 16u, table + r);
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    if (MODE == 0 || MODE == 2) {
      const int j0 = MODE == 2 ? per_thread / 2 : 0;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j0 + j < per_thread) v[j] = __ldg(table + (mix(seed + tid * 64u + j0 + j) & mask));
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j0 + j < per_thread) acc += v[j].x + v[j].w;
    }
    if (n_bulk) {
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (MODE != 0) acc += reinterpret_cast<float*>(smem)[(tid * 4 + it) & 1023];
    __syncthreads();
  }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, log2rows = argc > 2 ? atoi(argv[2]) : 22, threads = argc > 3 ? atoi(argv[3]) : 512;
  const int per_thread = argc > 4 ? atoi(argv[4]) : 16, iters = argc > 5 ? atoi(argv[5]) : 200;
  const size_t rows = (size_t)1 << log2rows;  // 16-byte rows
  float4* table;
  float* out;
  cudaMalloc(&table, rows * 16);
  cudaMemset(table, 0, rows * 16);
  cudaMalloc(&out, 4);
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = threads * per_thread * 16;
  auto launch = [&](int it) {
    switch (mode) {
      case 0: probe<0><<<sms, threads, smem>>>(table, (uint32_t)rows - 1, per_thread, it, out); break;
      case 1: probe<1><<<sms, threads, smem>>>(table, (uint32_t)rows - 1, per_thread, it, out); break;
      case 2: probe<2><<<sms, threads, smem>>>(table, (uint32_t)rows - 1, per_thread, it, out); break;
      default: probe<3><<<sms, threads, smem>>>(table, (uint32_t)rows - 1, per_thread, it, out); break;
    }
  };
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  launch(10);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  launch(iters);
  cudaEventRecord(b);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double copies = (double)sms * threads * per_thread * iters;
  printf("mode %d rows 2^%d (%.0f MB) threads %d x %d: %.3f ms, %.2f G rows/s, %.2f ns per row per SM, %.1f GB/s\n", mode, log2rows,
         rows * 16 / 1e6, threads, per_thread, ms, copies / ms * 1e-6, ms * 1e6 / (copies / sms), copies * 16 / ms * 1e-6);
  return 0;
}

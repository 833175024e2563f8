// Microbenchmark: what does one B200 SM sustain in scattered fp32 reductions (red.global.add.{f32,v2.f32,v4.f32}) into an
// L2-resident table, as a function of resident warps, vector width and address pattern?  (The table-gradient scatter of the
// backward kernels is bound by this rate)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/red_probe tools/r2/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// PATTERN 3,4,5: 2 / 4 / 8 adjacent lanes add into the same row (what same-cell samples of a ray do without merging).
// PATTERN 0: every lane a random row; 1: the warp hits 32 consecutive rows at a random base; 2: lanes in runs of 4 share a
// 64-byte neighbourhood (4 consecutive rows) -- roughly what neighbouring samples of a ray do at the fine levels
// ACTIVE: lanes of every warp that issue (the others are predicated off): is the cost per instruction or per active lane?
template <int VEC, int PATTERN, int ACTIVE = 32>
__global__ void red_kernel(float* table, uint32_t row_mask, int iters) {
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, gw = gt >> 5;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    uint32_t row;
    if (PATTERN == 0) row = mix(gt * 0x9E3779B9u + i);
    else if (PATTERN == 1) row = mix(gw * 0x9E3779B9u + i) + lane;
    else if (PATTERN == 2) row = mix((gt >> 2) * 0x9E3779B9u + i) * 4u + (lane & 3u);
    else row = mix((gt >> (PATTERN > 2 ? PATTERN - 2 : 0)) * 0x9E3779B9u + i);  // 3 / 4 / 5: groups of 2 / 4 / 8 adjacent lanes hit the SAME row
    row &= row_mask;
    if (ACTIVE < 32 && (lane % (32 / ACTIVE)) != 0) continue;
    float* p = table + (size_t)row * 2;  // rows of 2 floats, as the hash-table gradient
    if (VEC == 1) {
      atomicAdd(p, 1.0f);
    } else if (VEC == 2) {
      atomicAdd(reinterpret_cast<float2*>(p), make_float2(1.0f, 1.0f));
    } else {
      atomicAdd(reinterpret_cast<float4*>(table + (size_t)(row & ~1u) * 2), make_float4(1.0f, 1.0f, 1.0f, 1.0f));
    }
  }
}

template <int VEC, int PATTERN, int ACTIVE = 32>
void run(float* table, uint32_t rows, int sms, double ghz, bool quick = false) {
  for (int warps : {4, 8, 16, 32, 64}) {
    if (quick && warps != 4 && warps != 16) continue;
    const int threads = warps >= 32 ? 1024 : warps * 32;
    const int blocks_per_sm = warps >= 32 ? warps / 32 : 1;
    const int iters = 4096 / (warps >= 16 ? warps / 8 : 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    red_kernel<VEC, PATTERN, ACTIVE><<<sms * blocks_per_sm, threads>>>(table, rows - 1, iters / 4);  // warm-up
    cudaEventRecord(e0);
    red_kernel<VEC, PATTERN, ACTIVE><<<sms * blocks_per_sm, threads>>>(table, rows - 1, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lanes = (double)sms * blocks_per_sm * threads * iters * ACTIVE / 32.0;
    const double per_sm_cycle = lanes / sms / (ms * 1e-3 * ghz * 1e9);
    if (ACTIVE < 32) printf("[%2d of 32 lanes active: %.4f warp-instructions/clk/SM] ", ACTIVE, per_sm_cycle / ACTIVE);
    printf("vec %d pattern %d warps/SM %2d: %7.3f ms  %7.2f G lanes/s  %.3f lanes/clk/SM  %.2f B/clk/SM  (%s)\n", VEC, PATTERN, warps, ms,
           lanes / ms * 1e-6, per_sm_cycle, per_sm_cycle * 4 * VEC, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const double ghz = prop.clockRate * 1e-6;
  printf("%s: %d SMs, %.3f GHz nominal\n", prop.name, sms, ghz);
  for (uint32_t log2rows : {19u, 23u}) {  // 4 MB (one level of the T=19 table) and 64 MB (all 16 levels)
    const uint32_t rows = 1u << log2rows;
    float* table;
    cudaMalloc(&table, (size_t)rows * 2 * sizeof(float));
    cudaMemset(table, 0, (size_t)rows * 2 * sizeof(float));
    printf("--- table of 2^%u rows x 2 floats (%u MB)\n", log2rows, rows * 8u >> 20);
    run<1, 0>(table, rows, sms, ghz);
    run<2, 0>(table, rows, sms, ghz);
    run<4, 0>(table, rows, sms, ghz);
    run<2, 2>(table, rows, sms, ghz);
    run<4, 2>(table, rows, sms, ghz);
    run<2, 1>(table, rows, sms, ghz);
    run<2, 3>(table, rows, sms, ghz, true);
    run<2, 4>(table, rows, sms, ghz, true);
    run<2, 5>(table, rows, sms, ghz, true);
    run<2, 0, 16>(table, rows, sms, ghz, true);
    run<2, 0, 8>(table, rows, sms, ghz, true);
    run<2, 0, 4>(table, rows, sms, ghz, true);
    run<2, 0, 1>(table, rows, sms, ghz, true);
    run<4, 0, 16>(table, rows, sms, ghz, true);
    if (log2rows == 23u) {
      printf("--- same, on a half / a quarter of the SMs (is the limit per SM or chip-wide?)\n");
      run<2, 0>(table, rows, sms / 2, ghz, true);
      run<2, 0>(table, rows, sms / 4, ghz, true);
      run<4, 2>(table, rows, sms / 2, ghz, true);
    }
    cudaFree(table);
  }
  return 0;
}

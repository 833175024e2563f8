// Helpers shared by the fused tcgen05 forward (fnr_tc.cu) and backward (fnr_tc_bwd.cu) kernels.
#pragma once
#include "fnr_common.cuh"
#include "fnr_tcgen05.cuh"

namespace fnr {
namespace tcx {
using namespace tc;

constexpr unsigned kTcFullMask = 0xffffffffu;
constexpr int wbytes(int n, int k) { return n * k * 2; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Stage W[N][K] (fp32, torch layout, optional column permutation / row padding) into the canonical
// bf16 hi/lo tiles.  getw(n, k) returns the fp32 weight of padded position (n, k).
template <int NTHREADS, int NP, int KP, class F>
__device__ __forceinline__ void stage_weight(uint8_t* tile, F getw) {
  for (int idx = threadIdx.x; idx < NP * KP; idx += NTHREADS) {
    const int n = idx / KP, k = idx % KP;
    float hi, lo;
    split_bf16(getw(n, k), hi, lo);
    const int off = (k >> 3) * (NP * 16) + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16_rn(hi);
    *reinterpret_cast<__nv_bfloat16*>(tile + wbytes(NP, KP) + off) = __float2bfloat16_rn(lo);
  }
}

// Store 8 consecutive K elements of this thread's row (chunk j) as hi/lo bf16.
__device__ __forceinline__ void store_chunk(uint8_t* tile_hi, int lo_off, int row, int j, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], h[q], l[q]);
  uint8_t* p = tile_hi + j * (128 * 16) + row * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + lo_off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Issue the 3-way split GEMM D[128,N] = A[128,K] W[N,K]^T (one thread); a_lo = address of A's lo half.
template <int K, int N>
__device__ __forceinline__ void issue_gemm_lo(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi) {
  constexpr uint32_t idesc = idesc_bf16_f32(128, N);
  constexpr uint32_t w_lo_off = N * K * 2;
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t ah = smem_desc(a_hi + ks * 2 * 128 * 16, 128 * 16, 128);
    const uint64_t al = smem_desc(a_lo + ks * 2 * 128 * 16, 128 * 16, 128);
    const uint64_t wh = smem_desc(w_hi + ks * 2 * N * 16, N * 16, 128);
    const uint64_t wl = smem_desc(w_hi + w_lo_off + ks * 2 * N * 16, N * 16, 128);
    mma_ss(d_tmem, ah, wh, idesc, ks > 0);
    mma_ss(d_tmem, al, wh, idesc, true);
    mma_ss(d_tmem, ah, wl, idesc, true);
  }
}
template <int K, int N>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t w_hi) {
  issue_gemm_lo<K, N>(d_tmem, a_hi, a_hi + 128 * K * 2, w_hi);
}

__device__ __forceinline__ float warp_incl_scan_f(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kTcFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kTcFullMask, v, o);
  return v;
}


}  // namespace tcx
}  // namespace fnr

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


// The 8 corner contributions (v0[k], v1[k]) of one point at one level.  x-neighbour corners are adjacent table rows r, r^1 when the
// floor x is even: one 16-byte red for both (the cost of a red is per lane, not per byte: profiles/r2_red_probe.log); otherwise two
// 8-byte reds.  (Issuing 16-byte reds with a zero half for the unpaired lanes as well -- 8 instead of 12 red instructions per warp and
// level -- measured the same time.)
__device__ __forceinline__ void corner_reds(float2* gtab, const LevelCell& c, const float (&v0)[8], const float (&v1)[8], uint32_t hmask, uint32_t base) {
  const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
  constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};  // corner pairs (x floor, x ceil) per (y, z)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t rf = corner_row(c, kf[q], hmask, base);
    if (pair) {
      const bool f_first = (rf & 1u) == 0u;
      const float4 v = f_first ? make_float4(v0[kf[q]], v1[kf[q]], v0[kc[q]], v1[kc[q]]) : make_float4(v0[kc[q]], v1[kc[q]], v0[kf[q]], v1[kf[q]]);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) atomicAdd(reinterpret_cast<float4*>(gtab + (rf & ~1u)), v);
    } else {
      if (v0[kf[q]] != 0.f || v1[kf[q]] != 0.f) atomicAdd(gtab + rf, make_float2(v0[kf[q]], v1[kf[q]]));
      if (v0[kc[q]] != 0.f || v1[kc[q]] != 0.f) atomicAdd(gtab + corner_row(c, kc[q], hmask, base), make_float2(v0[kc[q]], v1[kc[q]]));
    }
  }
}

// ---- hash-table gradient of one level for a warp of points (lane = point; the 32 lanes are consecutive samples of a ray) ----------
// Coarse / middle levels: consecutive lanes share grid cells.  A run of lanes in the same cell is summed with a segmented suffix scan
// and only its head lane issues the reds (run length ~25 at level 0, ~9 at level 3, ~2 at level 8 on the bench batch); x-neighbour
// corners go out as one 16-byte red when they are adjacent table rows.  Reds are bound by the chip-wide L2 atomic REQUEST rate, so a
// merged run saves its requests whatever its width.
__device__ __forceinline__ void scatter_level_aggregated(float2* gtab, const Vec3& pos, bool live, float g0, float g1, int l, float scale, int log2T,
                                                         uint32_t hmask, int lane) {
  const LevelCell c = level_cell(pos, scale);
  const uint32_t key = live ? (c.hx[0] ^ c.hy[0] ^ c.hz[0]) : (0x80000000u | (uint32_t)lane);
  const uint32_t prev = __shfl_up_sync(kTcFullMask, key, 1);
  const bool head = lane == 0 || prev != key;
  const uint32_t heads = __ballot_sync(kTcFullMask, head);
  const uint32_t above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
  const int run_end = above ? (__ffs(above) - 1) : 32;
  bool same[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) same[q] = lane + (1 << q) < run_end;
  if (!live) g0 = g1 = 0.f;
  float v0[8], v1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = corner_weight(c, k);
    v0[k] = w * g0;
    v1[k] = w * g1;
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int dd = 1 << q;
    if (__any_sync(kTcFullMask, same[q])) {  // warp-uniform: skip the steps no run in this warp is long enough for
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t0 = __shfl_down_sync(kTcFullMask, v0[k], dd), t1 = __shfl_down_sync(kTcFullMask, v1[k], dd);
        if (same[q]) {
          v0[k] += t0;
          v1[k] += t1;
        }
      }
    }
  }
  if (head && live) {
    const uint32_t base = (uint32_t)l << log2T;
    const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
    constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};  // corner pairs (x floor, x ceil) per (y, z)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t rf = corner_row(c, kf[q], hmask, base);
      if (pair) {
        const bool f_first = (rf & 1u) == 0u;
        const float4 v = f_first ? make_float4(v0[kf[q]], v1[kf[q]], v0[kc[q]], v1[kc[q]]) : make_float4(v0[kc[q]], v1[kc[q]], v0[kf[q]], v1[kf[q]]);
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) atomicAdd(reinterpret_cast<float4*>(gtab + (rf & ~1u)), v);
      } else {
        if (v0[kf[q]] != 0.f || v1[kf[q]] != 0.f) atomicAdd(gtab + rf, make_float2(v0[kf[q]], v1[kf[q]]));
        if (v0[kc[q]] != 0.f || v1[kc[q]] != 0.f) atomicAdd(gtab + corner_row(c, kc[q], hmask, base), make_float2(v0[kc[q]], v1[kc[q]]));
      }
    }
  }
}
// One shuffle round instead of the full scan: the lane at an even position of a same-cell run absorbs its successor and the odd
// lanes stay silent -- half the red lanes of a long run for 16 shuffles (the full scan: 80).  (Two rounds -- blocks of four lanes --
// measured slower in the small backward: 0.806 against 0.791 ms, profiles/r2_backward_experiments.md.)
__device__ __forceinline__ void scatter_level_merged1(float2* gtab, const Vec3& pos, bool live, float g0, float g1, int l, float scale, int log2T,
                                                      uint32_t hmask, int lane) {
  const LevelCell c = level_cell(pos, scale);
  const uint32_t key = live ? (c.hx[0] ^ c.hy[0] ^ c.hz[0]) : (0x80000000u | (uint32_t)lane);
  const uint32_t prev = __shfl_up_sync(kTcFullMask, key, 1);
  const uint32_t heads = __ballot_sync(kTcFullMask, lane == 0 || prev != key);
  const int run_start = 31 - __clz(heads & ((2u << lane) - 1u));
  const bool odd = ((lane - run_start) & 1) != 0;  // absorbed by its predecessor
  const uint32_t odds = __ballot_sync(kTcFullMask, odd);
  const bool absorb = !odd && lane < 31 && ((odds >> (lane + 1)) & 1u);
  if (!live) g0 = g1 = 0.f;
  float v0[8], v1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = corner_weight(c, k);
    v0[k] = w * g0;
    v1[k] = w * g1;
  }
  if (odds != 0u) {  // warp-uniform
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t0 = __shfl_down_sync(kTcFullMask, v0[k], 1), t1 = __shfl_down_sync(kTcFullMask, v1[k], 1);
      if (absorb) {
        v0[k] += t0;
        v1[k] += t1;
      }
    }
  }
  if (live && !odd) corner_reds(gtab, c, v0, v1, hmask, (uint32_t)l << log2T);
}
// fine levels: x-neighbours (floor x even, ceil = floor + 1) are adjacent table rows: one 16-byte red instead of two 8-byte ones
__device__ __forceinline__ void scatter_level_direct(float2* gtab, const Vec3& pos, float g0, float g1, int l, float scale, int log2T, uint32_t hmask) {
  if (g0 == 0.f && g1 == 0.f) return;
  const LevelCell c = level_cell(pos, scale);
  const uint32_t base = (uint32_t)l << log2T;
  const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
  constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float wf = corner_weight(c, kf[q]), wc = corner_weight(c, kc[q]);
    const uint32_t rf = corner_row(c, kf[q], hmask, base);
    if (pair) {
      const uint32_t r0 = rf & ~1u;
      const bool f_first = (rf & 1u) == 0u;
      const float4 v = f_first ? make_float4(wf * g0, wf * g1, wc * g0, wc * g1) : make_float4(wc * g0, wc * g1, wf * g0, wf * g1);
      atomicAdd(reinterpret_cast<float4*>(gtab + r0), v);
    } else {
      if (wf != 0.f) atomicAdd(gtab + rf, make_float2(wf * g0, wf * g1));
      if (wc != 0.f) atomicAdd(gtab + corner_row(c, kc[q], hmask, base), make_float2(wc * g0, wc * g1));
    }
  }
}

}  // namespace tcx
}  // namespace fnr

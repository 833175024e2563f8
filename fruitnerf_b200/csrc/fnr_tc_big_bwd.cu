// Backward of the fruit_nerf_big family on the tensor cores, in two stages:
//   1. tc_big_backward_chain_kernel: per 128-point tile, recompute the activations from the stashed encoding (4 GEMM round
//      trips, bit-identical to fnr_tc_big.cu, ReLU masks kept as register bitmasks), then back-propagate dY through the
//      colour / semantic / base MLPs (dX = dY W with the forward weight tiles read MN-major, A operands in tensor memory),
//      scatter the encoding gradient into the hash-table gradient and the appearance-embedding gradient -- and write every
//      layer's input X (with a constant-1 column) and pre-activation gradient dY as fp32 row-major matrices to scratch.
//   2. the weight / bias gradients dW = dY^T [X | 1] are plain GEMMs with the 786 k points as reduction dimension: they are
//      handed to cuBLAS (TF32 tensor cores, resolved with dlopen so the library has no link-time dependency), followed by
//      one unpack kernel that adds the products into the torch-layout gradient tensors (column permutation of the colour
//      input, the folded semantic tail: d W_sem2 = head_w (x) v, d head_w = W_sem2 v + b_sem2 s, ...).
// The 128-wide semantic layers make TMEM-resident dW accumulators (fnr_tc_bwd.cu) impossible here: 416 accumulator columns
// plus the 384 columns of the chain exceed the 512 available.  Scratch: ~4.4 KB per point.
#include <cublas_v2.h>
#include <dlfcn.h>
#include <cstdlib>
#include "fnr_common.cuh"
#include "fnr_kernels.h"
#include "fnr_tcgen05.cuh"
#include "fnr_tc_common.cuh"

namespace fnr {
using namespace tc;
using namespace tcx;

namespace {

constexpr int kComputeThreads = 256;  // 8 warps: the tensor chain (thread pair (row, half) per point)
constexpr int kScatterThreads = 128;  // 4 warps: hash-table gradient reds of the previous tile, overlapped with the chain
                                      // (8 scatter warps measured slower: 3.32 vs 3.07 ms backward phase -- the reds contend)
constexpr int kCtaThreads = kComputeThreads + kScatterThreads;
constexpr int BAR_COMPUTE = 1, BAR_FULL = 2, BAR_EMPTY = 3;
constexpr int kAggLevels = 4;    // coarse levels whose reds are run-length aggregated across the warp
constexpr int STAGE_STRIDE = 36;  // floats per point in the hand-off buffer: denc[32], pos xyz, live flag
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int GEO = 30, ENC = 32, H = 64, APP = 32, SHD = 16, SW = 128, SOUT = 64, CIN = SHD + GEO + APP;
constexpr int K_BASE0 = 32, N_BASE0 = 64;
constexpr int K_BASE1 = 64, N_BASE1 = 32;
constexpr int K_SEM0 = 32, N_SEM0 = 128;
constexpr int K_SEM1 = 128, N_SEM1 = 128;
constexpr int K_COL0 = 80, N_COL0 = 64;
constexpr int K_COL1 = 64, N_COL1 = 64;
constexpr int K_COL2 = 64, N_COL2 = 16;

constexpr int OFF_W_BASE0 = 0;
constexpr int OFF_W_BASE1 = OFF_W_BASE0 + 2 * wbytes(N_BASE0, K_BASE0);
constexpr int OFF_W_SEM0 = OFF_W_BASE1 + 2 * wbytes(N_BASE1, K_BASE1);
constexpr int OFF_W_SEM1 = OFF_W_SEM0 + 2 * wbytes(N_SEM0, K_SEM0);
constexpr int OFF_W_COL0 = OFF_W_SEM1 + 2 * wbytes(N_SEM1, K_SEM1);
constexpr int OFF_W_COL1 = OFF_W_COL0 + 2 * wbytes(N_COL0, K_COL0);
constexpr int OFF_W_COL2 = OFF_W_COL1 + 2 * wbytes(N_COL1, K_COL1);
constexpr int OFF_F32 = OFF_W_COL2 + 2 * wbytes(N_COL2, K_COL2);
// floats: biases base0[64] base1[32] sem0[128] sem1[128] col0[64] col1[64] | fold[128]
constexpr int B_BASE0 = 0, B_BASE1 = 64, B_SEM0 = 96, B_SEM1 = 224, B_COL0 = 352, B_COL1 = 416, B_FOLD = 480, B_COUNT = 608;
constexpr int OFF_STAGE = OFF_F32 + B_COUNT * 4;
constexpr int kSmemBytes = OFF_STAGE + 128 * STAGE_STRIDE * 4 + 1024;
static_assert(OFF_STAGE % 16 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared-memory budget");

constexpr int R_A0 = 0, R_A1 = 128, R_D0 = 256, R_D1 = 384;

// scratch matrices (fp32, row-major, one row per point); X widths include the constant-1 column (+ padding to 4 floats)
// every width is a multiple of 8 floats so that each thread writes whole 32-byte sectors (st.global.v8.f32);
// layouts: XE [enc 32 | 1 | 0..], XH [h 64 | 1 | ..], XG [h0-slot, geo 30, pad | 1 | ..] (the kernel's K order),
// XZ [z 128 | 1 | ..], XC [sh 16 | app 32 | h0-slot, geo 30, pad | 1 | ..] (kernel K order; big_unpack_kernel permutes)
constexpr int XW_E = 40, XW_H = 72, XW_G = 40, XW_Z = 136, XW_C = 88, XW_C1 = 72, XW_C2 = 72;
constexpr int DW_H = 64, DW_OUT = 32, DW_Z = 128, DW_C = 64, DW_R = 8;
constexpr int kFloatsPerPoint = XW_E + XW_H + XW_G + 2 * XW_Z + XW_C + XW_C1 + XW_C2 + DW_H + DW_OUT + 2 * DW_Z + 2 * DW_C + DW_R;
// GEMM outputs C_l [N_l x XW_l]
constexpr int CO_B0 = 0, CO_B1 = CO_B0 + 64 * XW_E, CO_S0 = CO_B1 + 32 * XW_H, CO_S1 = CO_S0 + 128 * XW_G, CO_F = CO_S1 + 128 * XW_Z,
              CO_C0 = CO_F + 4 * XW_Z, CO_C1 = CO_C0 + 64 * XW_C, CO_C2 = CO_C1 + 64 * XW_C1, CO_END = CO_C2 + 4 * XW_C2;

struct Bufs {
  float *xe, *xh, *xg, *xz1, *xz2, *xc, *xc1, *xc2;
  float *dh, *dout, *dz1, *dz2, *dc1, *dc2, *dr;
  float* cout;  // GEMM outputs [CO_END]
};

struct ChainArgs {
  KField F;
  KParams P;
  KParams G;
  KRays Rr;
  const float* point_grads;
  const float* stash;
  const float* sample_rgb;
  Bufs B;
  int debug_flags;  // FNR_DEBUG_BWD (timing experiments only): bit 0 skip the table scatter, bit 2 skip the X / dY stores
};

template <int K>
__device__ __forceinline__ void st_a16(uint32_t ab, int k0, const float (&v)[16]) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], h[q], l[q]);
  tmem_st8(ab + (k0 >> 1), h);
  tmem_st8(ab + K / 2 + (k0 >> 1), l);
}

// 256-bit stores (STG.E.ENL2.256): one whole 32-byte sector per lane and instruction; dst must be 32-byte aligned
__device__ __forceinline__ void store8(float* dst, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6),
               "f"(a7)
               : "memory");
}
__device__ __forceinline__ void store16(float* dst, const float (&v)[16]) {
  store8(dst, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
  store8(dst + 8, v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
}
__device__ __forceinline__ void store_one_col(float* dst) { store8(dst, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); }

// forward GEMM: A (TMEM, depth K) x W[N,K]^T
template <int K, int N>
__device__ __forceinline__ void issue_fwd(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_hi) {
  constexpr uint32_t idesc = idesc_bf16_f32(128, N);
  constexpr uint32_t w_lo_off = N * K * 2;
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t wh = smem_desc(w_hi + ks * 2 * N * 16, N * 16, 128);
    const uint64_t wl = smem_desc(w_hi + w_lo_off + ks * 2 * N * 16, N * 16, 128);
    mma_ts(d_tmem, a_tmem + 8 * ks, wh, idesc, ks > 0);
    mma_ts(d_tmem, a_tmem + K / 2 + 8 * ks, wh, idesc, true);
    mma_ts(d_tmem, a_tmem + 8 * ks, wl, idesc, true);
  }
}

// dX[128, NOUT] = dY[128, KR] (TMEM, depth KR) x W[KR rows, NOUT cols]: the forward tile of W (KR = its N, NOUT = its K) read MN-major
template <int KR, int NOUT>
__device__ __forceinline__ void issue_dx(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_hi) {
  constexpr uint32_t idesc = idesc_bf16_f32(128, NOUT) | (1u << 16);  // b_major = MN
  constexpr uint32_t w_lo = KR * NOUT * 2;
#pragma unroll
  for (int ks = 0; ks < KR / 16; ++ks) {
    const uint64_t wh = smem_desc(w_hi + ks * 256, 128, KR * 16);
    const uint64_t wl = smem_desc(w_hi + w_lo + ks * 256, 128, KR * 16);
    mma_ts(d_tmem, a_tmem + 8 * ks, wh, idesc, ks > 0);
    mma_ts(d_tmem, a_tmem + KR / 2 + 8 * ks, wh, idesc, true);
    mma_ts(d_tmem, a_tmem + 8 * ks, wl, idesc, true);
  }
}

__global__ void __launch_bounds__(kCtaThreads, 1) tc_big_backward_chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_compute = tid < kComputeThreads;
  const int half = (warp >> 2) & 1;
  const int row = (warp & 3) * 32 + lane;  // compute: TMEM lane / point of the tile; scatter warps 8..11: the point they scatter
  const KParams& P = a.P;
  const KParams& G = a.G;
  const KField& F = a.F;
  const Bufs& B = a.B;
  float* sf = reinterpret_cast<float*>(smem + OFF_F32);
  float* stage = reinterpret_cast<float*>(smem + OFF_STAGE);

  if (warp == 0) tmem_alloc(&s_tmem_base, 512);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  stage_weight<kCtaThreads, N_BASE0, K_BASE0>(smem + OFF_W_BASE0, [&](int n, int k) { return __ldg(P.base_w[0] + n * ENC + k); });
  stage_weight<kCtaThreads, N_BASE1, K_BASE1>(smem + OFF_W_BASE1, [&](int n, int k) { return n < 1 + GEO ? __ldg(P.base_w[1] + n * H + k) : 0.f; });
  stage_weight<kCtaThreads, N_SEM0, K_SEM0>(smem + OFF_W_SEM0,
                                            [&](int n, int k) { return (k >= 1 && k <= GEO) ? __ldg(P.sem_w[0] + n * GEO + (k - 1)) : 0.f; });
  stage_weight<kCtaThreads, N_SEM1, K_SEM1>(smem + OFF_W_SEM1, [&](int n, int k) { return __ldg(P.sem_w[1] + n * SW + k); });
  stage_weight<kCtaThreads, N_COL0, K_COL0>(smem + OFF_W_COL0, [&](int n, int k) {
    const float* w = P.col_w[0] + n * CIN;
    if (k < SHD) return __ldg(w + k);
    if (k < SHD + APP) return __ldg(w + SHD + GEO + (k - SHD));
    const int g = k - SHD - APP - 1;
    return (g >= 0 && g < GEO) ? __ldg(w + SHD + g) : 0.f;
  });
  stage_weight<kCtaThreads, N_COL1, K_COL1>(smem + OFF_W_COL1, [&](int n, int k) { return __ldg(P.col_w[1] + n * H + k); });
  stage_weight<kCtaThreads, N_COL2, K_COL2>(smem + OFF_W_COL2, [&](int n, int k) { return n < 3 ? __ldg(P.col_w[2] + n * H + k) : 0.f; });
  for (int i = tid; i < B_COUNT; i += kCtaThreads) {
    float v = 0.f;
    if (i < B_BASE1) v = __ldg(P.base_b[0] + i);
    else if (i < B_SEM0) v = (i - B_BASE1) < 1 + GEO ? __ldg(P.base_b[1] + (i - B_BASE1)) : 0.f;
    else if (i < B_SEM1) v = __ldg(P.sem_b[0] + (i - B_SEM0));
    else if (i < B_COL0) v = __ldg(P.sem_b[1] + (i - B_SEM1));
    else if (i < B_COL1) v = __ldg(P.col_b[0] + (i - B_COL0));
    else if (i < B_FOLD) v = __ldg(P.col_b[1] + (i - B_COL1));
    else {  // fold[k] = sum_j head_w[j] * W_sem2[j][k]: d logit / d z2[k]
      const int k = i - B_FOLD;
      float acc = 0.f;
      for (int j = 0; j < SOUT; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_w[2] + j * SW + k), acc);
      v = acc;
    }
    sf[i] = v;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = s_tmem_base;
  const uint32_t tr = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t wBase = smem_u32(smem);
  uint32_t phase = 0;

  const long long N = (long long)a.Rr.R * a.Rr.S;
  const long long tiles = (N + 127) / 128;
  const int S = a.Rr.S;
  const uint32_t hmask = (1u << F.log2T) - 1u;
  float2* const gtab = reinterpret_cast<float2*>(G.hash_table);

#define FNR_ISSUE(...)                          \
  tmem_st_wait();                               \
  fence_before_sync();                          \
  named_bar_sync(BAR_COMPUTE, kComputeThreads); \
  if (warp == 0) {                              \
    if (elect_one_sync()) {     \
      fence_after_sync();       \
      __VA_ARGS__;              \
      mma_commit(&s_bar);       \
    }                           \
    __syncwarp();               \
  }
#define FNR_WAIT()          \
  mbar_wait(&s_bar, phase); \
  phase ^= 1;               \
  fence_after_sync();

  if (!is_compute) {
    // ================= scatter warps: hash-table gradient reds, decoupled from the tensor chain =================
    reg_dec<72>();
    const bool do_scatter = !(a.debug_flags & 1);
    named_bar_arrive(BAR_EMPTY, kCtaThreads);
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      named_bar_sync(BAR_FULL, kCtaThreads);
      float g[32];
      const float4* src = reinterpret_cast<const float4*>(stage + row * STAGE_STRIDE);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = src[q];
        g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
      }
      const float4 pv = src[8];
      if (tile + gridDim.x < tiles) named_bar_arrive(BAR_EMPTY, kCtaThreads);  // the hand-off buffer may be overwritten
      const Vec3 pos = {pv.x, pv.y, pv.z};
      const bool live = pv.w != 0.f;
      if (do_scatter) {
        // Levels 0..kAggLevels-1 (as in fnr_tc_bwd.cu): the 32 lanes of a warp are consecutive samples of a ray and share grid
        // cells; a run of lanes in the same cell is summed with a segmented suffix scan and only its head lane issues the reds.
#pragma unroll 1
        for (int l = 0; l < kAggLevels; ++l) {
          const LevelCell c = level_cell(pos, F.scalings[l]);
          const uint32_t key = live ? (c.hx[0] ^ c.hy[0] ^ c.hz[0]) : (0x80000000u | (uint32_t)lane);
          const uint32_t prev = __shfl_up_sync(kTcFullMask, key, 1);
          const bool head = lane == 0 || prev != key;
          const uint32_t heads = __ballot_sync(kTcFullMask, head);
          const uint32_t above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
          const int run_end = above ? (__ffs(above) - 1) : 32;
          bool same[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) same[q] = lane + (1 << q) < run_end;
          float g0 = 0.f, g1 = 0.f;
#pragma unroll
          for (int ll = 0; ll < kAggLevels; ++ll)
            if (ll == l) {
              g0 = live ? g[2 * ll] : 0.f;
              g1 = live ? g[2 * ll + 1] : 0.f;
            }
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
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float t0 = __shfl_down_sync(kTcFullMask, v0[k], dd), t1 = __shfl_down_sync(kTcFullMask, v1[k], dd);
              if (same[q]) {
                v0[k] += t0;
                v1[k] += t1;
              }
            }
          }
          if (head && live) {
            const uint32_t base = (uint32_t)l << F.log2T;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (v0[k] != 0.f || v1[k] != 0.f) atomicAdd(gtab + corner_row(c, k, hmask, base), make_float2(v0[k], v1[k]));
          }
        }
      }
      if (live && do_scatter) {
#pragma unroll
        for (int l = kAggLevels; l < 16; ++l) {
          const float g0 = g[2 * l], g1 = g[2 * l + 1];
          if (g0 != 0.f || g1 != 0.f) {
            const LevelCell c = level_cell(pos, F.scalings[l]);
            const uint32_t base = (uint32_t)l << F.log2T;
            const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);  // x-neighbours: one 16-byte red
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
        }
      }
    }
  } else {
  // ================= compute warps =================
  reg_inc<216>();
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long p = tile * 128 + row;
    const bool in_range = p < N;
    const bool valid = in_range && !(a.debug_flags & 4);  // `valid` guards the scratch stores
    const long long pc = in_range ? p : N - 1;
    const int ray = (int)(pc / S);
    const float* o = a.Rr.origins + 3 * (size_t)ray;
    const float* d = a.Rr.directions + 3 * (size_t)ray;
    bool sel;
    const Vec3 pos = field_position(o, d, __ldg(a.Rr.starts + pc), __ldg(a.Rr.ends + pc), F.position_mode, F.aabb, sel);
    const float vm = in_range ? 1.f : 0.f;
    const float* pg = a.point_grads + 5 * (size_t)pc;
    const float d_sigma = __ldg(pg) * vm;
    const float d_logit = __ldg(pg + 4) * vm;
    const int cam = (F.appearance_mode == FNR_APP_PER_CAMERA) ? __ldg(a.Rr.camera_indices + ray) : 0;

    // ---- R1: encoding (stash) -> A0 ; X_enc ----
    {
      float enc[16];
      const float4* s4 = reinterpret_cast<const float4*>(a.stash + (size_t)pc * ENC + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 u = __ldg(s4 + q);
        enc[4 * q] = u.x; enc[4 * q + 1] = u.y; enc[4 * q + 2] = u.z; enc[4 * q + 3] = u.w;
      }
      st_a16<K_BASE0>(tr + R_A0, 16 * half, enc);
      if (valid) {
        store16(B.xe + (size_t)p * XW_E + 16 * half, enc);
        if (half == 1) store_one_col(B.xe + (size_t)p * XW_E + 32);
      }
    }
    FNR_ISSUE(issue_fwd<K_BASE0, N_BASE0>(tb + R_D0, tb + R_A0, wBase + OFF_W_BASE0))

    // ---- epi1: h = relu(base0 + b) -> A1, mask, X_h ; sh / app blocks of the colour input -> A0, X_cin ----
    FNR_WAIT()
    uint32_t mask_h = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        v[q] = fmaxf(__uint_as_float(r[q]) + sf[B_BASE0 + 32 * half + 16 * j + q], 0.f);
        if (v[q] > 0.f) mask_h |= 1u << (16 * j + q);
      }
      st_a16<K_BASE1>(tr + R_A1, 32 * half + 16 * j, v);
      if (valid) store16(B.xh + (size_t)p * XW_H + 32 * half + 16 * j, v);
    }
    if (valid && half == 1) store_one_col(B.xh + (size_t)p * XW_H + 64);
    if (half == 0) {
      float sh[16];
      sh_degree4(__ldg(d), __ldg(d + 1), __ldg(d + 2), sh);
      st_a16<K_COL0>(tr + R_A0, 0, sh);
      if (valid) store16(B.xc + (size_t)p * XW_C, sh);
    } else {
      const float* app = (F.appearance_mode == FNR_APP_PER_CAMERA) ? P.app_embedding + (size_t)cam * APP : nullptr;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = app ? __ldg(app + 16 * j + q) : 0.f;
        st_a16<K_COL0>(tr + R_A0, SHD + 16 * j, v);
        if (valid) store16(B.xc + (size_t)p * XW_C + SHD + 16 * j, v);
      }
      if (valid) store_one_col(B.xc + (size_t)p * XW_C + K_COL0);
    }
    FNR_ISSUE(issue_fwd<K_BASE1, N_BASE1>(tb + R_D1, tb + R_A1, wBase + OFF_W_BASE1))

    // ---- epi2: [h0 | geo | pad] -> semantic input (A1 + 64) and the last block of the colour input ; X_geo, X_cin ----
    FNR_WAIT()
    float dsig_scale = 0.f;
    {
      uint32_t r0[16];
      tmem_ld16(tr + R_D1 + 16 * half, r0);
      tmem_ld_wait();
      float g[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) g[q] = __uint_as_float(r0[q]) + sf[B_BASE1 + 16 * half + q];
      if (half == 0) dsig_scale = sel ? expf(fminf(fmaxf(g[0], -15.f), 15.f)) : 0.f;  // trunc_exp backward
      st_a16<K_SEM0>(tr + R_A1 + 64, 16 * half, g);
      st_a16<K_COL0>(tr + R_A0, SHD + APP + 16 * half, g);
      if (valid) {
        store16(B.xg + (size_t)p * XW_G + 16 * half, g);
        store16(B.xc + (size_t)p * XW_C + SHD + APP + 16 * half, g);
        if (half == 1) store_one_col(B.xg + (size_t)p * XW_G + 32);
      }
    }
    FNR_ISSUE(issue_fwd<K_SEM0, N_SEM0>(tb + R_D0, tb + R_A1 + 64, wBase + OFF_W_SEM0);
              issue_fwd<K_COL0, N_COL0>(tb + R_D1 + 64, tb + R_A0, wBase + OFF_W_COL0))

    // ---- epi3: z1 -> A1 (K 128), c1 -> A0 (K 64) ; masks ; X_z1, X_c1 ----
    FNR_WAIT()
    uint32_t mask_z1[2] = {0u, 0u}, mask_c1 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 64 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        v[q] = fmaxf(__uint_as_float(r[q]) + sf[B_SEM0 + 64 * half + 16 * j + q], 0.f);
        if (v[q] > 0.f) mask_z1[j >> 1] |= 1u << (16 * (j & 1) + q);
      }
      st_a16<K_SEM1>(tr + R_A1, 64 * half + 16 * j, v);
      if (valid) store16(B.xz1 + (size_t)p * XW_Z + 64 * half + 16 * j, v);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 64 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        v[q] = fmaxf(__uint_as_float(r[q]) + sf[B_COL0 + 32 * half + 16 * j + q], 0.f);
        if (v[q] > 0.f) mask_c1 |= 1u << (16 * j + q);
      }
      st_a16<K_COL1>(tr + R_A0, 32 * half + 16 * j, v);
      if (valid) store16(B.xc1 + (size_t)p * XW_C1 + 32 * half + 16 * j, v);
    }
    if (valid && half == 1) {
      store_one_col(B.xz1 + (size_t)p * XW_Z + 128);
      store_one_col(B.xc1 + (size_t)p * XW_C1 + 64);
    }
    FNR_ISSUE(issue_fwd<K_SEM1, N_SEM1>(tb + R_D0, tb + R_A1, wBase + OFF_W_SEM1);
              issue_fwd<K_COL1, N_COL1>(tb + R_D1, tb + R_A0, wBase + OFF_W_COL1))

    // ---- epi4: z2, c2 (X_z2, X_c2, masks) ; dz2 = dlogit * fold * relu'(z2) -> A1 ; d rgb_pre -> A0 (K 16) ----
    FNR_WAIT()
    uint32_t mask_c2 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 64 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16], dz[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int n = 64 * half + 16 * j + q;
        v[q] = fmaxf(__uint_as_float(r[q]) + sf[B_SEM1 + n], 0.f);
        dz[q] = v[q] > 0.f ? d_logit * sf[B_FOLD + n] : 0.f;
      }
      st_a16<K_SEM1>(tr + R_A1, 64 * half + 16 * j, dz);
      if (valid) {
        store16(B.xz2 + (size_t)p * XW_Z + 64 * half + 16 * j, v);
        store16(B.dz2 + (size_t)p * DW_Z + 64 * half + 16 * j, dz);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        v[q] = fmaxf(__uint_as_float(r[q]) + sf[B_COL1 + 32 * half + 16 * j + q], 0.f);
        if (v[q] > 0.f) mask_c2 |= 1u << (16 * j + q);
      }
      if (valid) store16(B.xc2 + (size_t)p * XW_C2 + 32 * half + 16 * j, v);
    }
    if (valid && half == 1) {
      store_one_col(B.xz2 + (size_t)p * XW_Z + 128);
      store_one_col(B.xc2 + (size_t)p * XW_C2 + 64);
    }
    if (half == 0) {
      float dr[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) dr[q] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float rgb = __ldg(a.sample_rgb + 3 * (size_t)pc + c);
        dr[c] = __ldg(pg + 1 + c) * vm * rgb * (1.0f - rgb);  // sigmoid'
      }
      st_a16<16>(tr + R_A0, 0, dr);
      if (valid) store8(B.dr + (size_t)p * DW_R, dr[0], dr[1], dr[2], 0.f, d_logit, 0.f, 0.f, 0.f);  // columns 4..7: dY of the folded tail
    }
    FNR_ISSUE(issue_dx<N_SEM1, K_SEM1>(tb + R_D0, tb + R_A1, wBase + OFF_W_SEM1);
              issue_dx<N_COL2, K_COL2>(tb + R_D1, tb + R_A0, wBase + OFF_W_COL2))

    // ---- epi6: dz1 = (dz2 W_sem1) * relu'(z1) -> dY only (semantic branch sees detach(geo)) ; dc2 -> A0 ----
    FNR_WAIT()
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 64 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = ((mask_z1[j >> 1] >> (16 * (j & 1) + q)) & 1u) ? __uint_as_float(r[q]) : 0.f;
      if (valid) store16(B.dz1 + (size_t)p * DW_Z + 64 * half + 16 * j, v);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = ((mask_c2 >> (16 * j + q)) & 1u) ? __uint_as_float(r[q]) : 0.f;
      st_a16<K_COL2>(tr + R_A0, 32 * half + 16 * j, v);
      if (valid) store16(B.dc2 + (size_t)p * DW_C + 32 * half + 16 * j, v);
    }
    FNR_ISSUE(issue_dx<N_COL1, K_COL1>(tb + R_D1 + 64, tb + R_A0, wBase + OFF_W_COL1))

    // ---- epi7: dc1 = (dc2 W_col1) * relu'(c1) -> A0 ----
    FNR_WAIT()
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 64 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = ((mask_c1 >> (16 * j + q)) & 1u) ? __uint_as_float(r[q]) : 0.f;
      st_a16<K_COL1>(tr + R_A0, 32 * half + 16 * j, v);
      if (valid) store16(B.dc1 + (size_t)p * DW_C + 32 * half + 16 * j, v);
    }
    FNR_ISSUE(issue_dx<N_COL0, K_COL0>(tb + R_D0, tb + R_A0, wBase + OFF_W_COL0))

    // ---- epi8: d cin: appearance-embedding gradient ; d[h0 | geo | pad] (+ d sigma) -> A1 (K 32), dY_out ----
    FNR_WAIT()
    {
      uint32_t ra[16], rg[16];
      tmem_ld16(tr + R_D0 + SHD + 16 * half, ra);         // d app[16 half .. +16)
      tmem_ld16(tr + R_D0 + SHD + APP + 16 * half, rg);   // d [h0-slot | geo | pad][16 half .. +16)
      tmem_ld_wait();
      if (F.appearance_mode == FNR_APP_PER_CAMERA) {
        const bool uniform = __all_sync(kTcFullMask, cam == __shfl_sync(kTcFullMask, cam, 0));
        if (uniform) {  // transposed butterfly: lane L ends with the (half-)sum of value L >> 1
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = in_range ? __uint_as_float(ra[i]) : 0.f;
#pragma unroll
          for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < n) {
                const float send = hi ? w[i] : w[i + n];
                const float keep = hi ? w[i + n] : w[i];
                w[i] = keep + __shfl_xor_sync(kTcFullMask, send, off);
              }
            }
          }
          const float tot = w[0] + __shfl_xor_sync(kTcFullMask, w[0], 1);
          if ((lane & 1) == 0 && tot != 0.f) atomicAdd(G.app_embedding + (size_t)cam * APP + 16 * half + (lane >> 1), tot);
        } else if (in_range) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = __uint_as_float(ra[i]);
            if (v != 0.f) atomicAdd(G.app_embedding + (size_t)cam * APP + 16 * half + i, v);
          }
        }
      }
      float dout[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) dout[q] = __uint_as_float(rg[q]);
      if (half == 0) dout[0] = d_sigma * dsig_scale;  // the h0 slot of the colour input carries no weight: its dX is 0
      else dout[15] = 0.f;                            // pad position
      st_a16<N_BASE1>(tr + R_A1, 16 * half, dout);
      if (valid) store16(B.dout + (size_t)p * DW_OUT + 16 * half, dout);
    }
    FNR_ISSUE(issue_dx<N_BASE1, K_BASE1>(tb + R_D1, tb + R_A1, wBase + OFF_W_BASE1))

    // ---- epi9: dh = (dout W_base1) * relu'(h) -> A0 (K 64), dY_h ----
    FNR_WAIT()
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = ((mask_h >> (16 * j + q)) & 1u) ? __uint_as_float(r[q]) : 0.f;
      st_a16<N_BASE0>(tr + R_A0, 32 * half + 16 * j, v);
      if (valid) store16(B.dh + (size_t)p * DW_H + 32 * half + 16 * j, v);
    }
    FNR_ISSUE(issue_dx<N_BASE0, K_BASE0>(tb + R_D0, tb + R_A0, wBase + OFF_W_BASE0))

    // ---- epi10: d enc -> hand-off to the scatter warps ----
    FNR_WAIT()
    {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 16 * half, r);
      tmem_ld_wait();
      named_bar_sync(BAR_EMPTY, kCtaThreads);  // the scatter warps have copied the previous tile out
      float4* dst = reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
      if (half == 0) reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 32)[0] = make_float4(pos.x, pos.y, pos.z, in_range ? 1.f : 0.f);
    }
    named_bar_arrive(BAR_FULL, kCtaThreads);
    fence_before_sync();
  }
  }  // compute warps
#undef FNR_ISSUE
#undef FNR_WAIT

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

// GEMM products -> torch-layout gradient tensors (+=)
__global__ void __launch_bounds__(256) big_unpack_kernel(const float* __restrict__ c, KParams P, KParams G) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = t; i < 64 * XW_E; i += nt) {
    const int n = i / XW_E, k = i % XW_E;
    const float v = c[CO_B0 + i];
    if (k < 32) G.base_w[0][n * ENC + k] += v;
    else if (k == 32) G.base_b[0][n] += v;
  }
  for (int i = t; i < 32 * XW_H; i += nt) {
    const int n = i / XW_H, k = i % XW_H;
    if (n >= 1 + GEO) continue;
    const float v = c[CO_B1 + i];
    if (k < 64) G.base_w[1][n * H + k] += v;
    else if (k == 64) G.base_b[1][n] += v;
  }
  for (int i = t; i < 128 * XW_G; i += nt) {
    const int n = i / XW_G, k = i % XW_G;
    const float v = c[CO_S0 + i];
    if (k >= 1 && k <= GEO) G.sem_w[0][n * GEO + (k - 1)] += v;  // column 0 is the h0 slot
    else if (k == 32) G.sem_b[0][n] += v;
  }
  for (int i = t; i < 128 * XW_Z; i += nt) {
    const int n = i / XW_Z, k = i % XW_Z;
    const float v = c[CO_S1 + i];
    if (k < 128) G.sem_w[1][n * SW + k] += v;
    else if (k == 128) G.sem_b[1][n] += v;
  }
  {  // folded tail: v[k] = sum dlogit z2[k], s = sum dlogit
    const float* vv = c + CO_F;
    const float s = vv[128];
    for (int i = t; i < SOUT * SW; i += nt) {
      const int j = i / SW, k = i % SW;
      G.sem_w[2][i] += P.head_w[j] * vv[k];
    }
    for (int j = t; j < SOUT; j += nt) {
      G.sem_b[2][j] += P.head_w[j] * s;
      float acc = P.sem_b[2][j] * s;
      for (int k = 0; k < SW; ++k) acc = fmaf(P.sem_w[2][j * SW + k], vv[k], acc);
      G.head_w[j] += acc;
    }
    if (t == 0) G.head_b[0] += s;
  }
  for (int i = t; i < 64 * XW_C; i += nt) {
    const int n = i / XW_C, k = i % XW_C;
    const float v = c[CO_C0 + i];  // kernel K order [sh | app | h0-slot, geo, pad | 1] -> torch order [sh | geo | app]
    if (k < SHD) G.col_w[0][n * CIN + k] += v;
    else if (k < SHD + APP) G.col_w[0][n * CIN + SHD + GEO + (k - SHD)] += v;
    else if (k >= SHD + APP + 1 && k <= SHD + APP + GEO) G.col_w[0][n * CIN + SHD + (k - SHD - APP - 1)] += v;
    else if (k == K_COL0) G.col_b[0][n] += v;
  }
  for (int i = t; i < 64 * XW_C1; i += nt) {
    const int n = i / XW_C1, k = i % XW_C1;
    const float v = c[CO_C1 + i];
    if (k < 64) G.col_w[1][n * H + k] += v;
    else if (k == 64) G.col_b[1][n] += v;
  }
  for (int i = t; i < 3 * XW_C2; i += nt) {
    const int n = i / XW_C2, k = i % XW_C2;
    const float v = c[CO_C2 + i];
    if (k < 64) G.col_w[2][n * H + k] += v;
    else if (k == 64) G.col_b[2][n] += v;
  }
}

// ---- cuBLAS, resolved at first use (no link-time dependency: the C-ABI library must load without it) ----
struct Cublas {
  void* lib = nullptr;
  cublasHandle_t handle = nullptr;
  cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*gemm_ex)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void*, const void*, cudaDataType, int,
                            const void*, cudaDataType, int, const void*, void*, cudaDataType, int, cublasComputeType_t, cublasGemmAlgo_t) = nullptr;
  bool ok = false;
};

Cublas& cublas() {
  static Cublas c;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12"}) {
      c.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (c.lib) break;
    }
    if (c.lib) {
      c.create = reinterpret_cast<decltype(c.create)>(dlsym(c.lib, "cublasCreate_v2"));
      c.set_stream = reinterpret_cast<decltype(c.set_stream)>(dlsym(c.lib, "cublasSetStream_v2"));
      c.gemm_ex = reinterpret_cast<decltype(c.gemm_ex)>(dlsym(c.lib, "cublasGemmEx"));
      if (c.create && c.set_stream && c.gemm_ex && c.create(&c.handle) == CUBLAS_STATUS_SUCCESS) c.ok = true;
    }
  }
  return c;
}

// row-major C[n_out x xw] = dY[P x dw (lda)]^T X[P x xw]
int gemm_dw(Cublas& cb, const float* dy, int ldd, int n_out, const float* x, int xw, long long Pn, float* cout) {
  const float one = 1.f, zero = 0.f;
  const cublasStatus_t s = cb.gemm_ex(cb.handle, CUBLAS_OP_N, CUBLAS_OP_T, xw, n_out, (int)Pn, &one, x, CUDA_R_32F, xw, dy, CUDA_R_32F, ldd, &zero,
                                      cout, CUDA_R_32F, xw, CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_GEMM_DEFAULT);
  if (s != CUBLAS_STATUS_SUCCESS) {
    set_error("cublasGemmEx failed with status %d", (int)s);
    return FNR_ERR_CUDA;
  }
  return FNR_OK;
}

}  // namespace

size_t tc_big_backward_scratch_bytes(long long num_points) {
  return (size_t)num_points * kFloatsPerPoint * sizeof(float) + (size_t)CO_END * sizeof(float) + 4096;
}

bool tc_big_backward_supported(const KField& F, const KFieldBwd& B) {
  return B.stash_encoding != nullptr && B.sample_rgb != nullptr && B.extra != nullptr && !F.pass_semantic_gradients &&
         (F.appearance_mode == FNR_APP_PER_CAMERA || F.appearance_mode == FNR_APP_ZEROS) && cublas().ok;
}

int launch_tc_big_field_backward(const KField& F, const KParams& P, const KParams& G, const KRays& Rr, const KFieldBwd& Bw, cudaStream_t st) {
  const long long N = (long long)Rr.R * Rr.S;
  if (N == 0) return FNR_OK;
  if (N > 0x7fffffffLL) {
    set_error("too many points for the cuBLAS reduction dimension");
    return FNR_ERR_UNSUPPORTED;
  }
  Cublas& cb = cublas();
  if (!cb.ok) {
    set_error("cuBLAS is not available (dlopen libcublas.so.12)");
    return FNR_ERR_UNSUPPORTED;
  }
  if (Bw.extra_bytes < tc_big_backward_scratch_bytes(N)) {
    set_error("scratch too small for the big-family tensor-core backward");
    return FNR_ERR_INVALID_ARGUMENT;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_big_backward_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_big_backward_chain_kernel)");
    configured = true;
  }
  ChainArgs a;
  {
    const char* e = getenv("FNR_DEBUG_BWD");
    a.debug_flags = e ? atoi(e) : 0;
  }
  a.F = F;
  a.P = P;
  a.G = G;
  a.Rr = Rr;
  a.point_grads = Bw.point_grads;
  a.stash = Bw.stash_encoding;
  a.sample_rgb = Bw.sample_rgb;
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Bw.extra) + 255) & ~(uintptr_t)255);
  auto take = [&](int width) {
    float* ptr = base;
    base += (size_t)N * width;
    return ptr;
  };
  Bufs& B = a.B;
  B.xe = take(XW_E); B.xh = take(XW_H); B.xg = take(XW_G); B.xz1 = take(XW_Z); B.xz2 = take(XW_Z); B.xc = take(XW_C); B.xc1 = take(XW_C1);
  B.xc2 = take(XW_C2); B.dh = take(DW_H); B.dout = take(DW_OUT); B.dz1 = take(DW_Z); B.dz2 = take(DW_Z); B.dc1 = take(DW_C); B.dc2 = take(DW_C);
  B.dr = take(DW_R);
  B.cout = base;
  const long long tiles = (N + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc_big_backward_chain_kernel<<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  if (int rc = check_cuda(cudaGetLastError(), "tc_big_backward_chain_kernel")) return rc;

  if (cb.set_stream(cb.handle, st) != CUBLAS_STATUS_SUCCESS) {
    set_error("cublasSetStream failed");
    return FNR_ERR_CUDA;
  }
  int rc;
  if ((rc = gemm_dw(cb, B.dh, DW_H, 64, B.xe, XW_E, N, B.cout + CO_B0))) return rc;
  if ((rc = gemm_dw(cb, B.dout, DW_OUT, 32, B.xh, XW_H, N, B.cout + CO_B1))) return rc;
  if ((rc = gemm_dw(cb, B.dz1, DW_Z, 128, B.xg, XW_G, N, B.cout + CO_S0))) return rc;
  if ((rc = gemm_dw(cb, B.dz2, DW_Z, 128, B.xz1, XW_Z, N, B.cout + CO_S1))) return rc;
  if ((rc = gemm_dw(cb, B.dr + 4, DW_R, 4, B.xz2, XW_Z, N, B.cout + CO_F))) return rc;  // dY = [d logit, 0, 0, 0]: row 0 = v | s
  if ((rc = gemm_dw(cb, B.dc1, DW_C, 64, B.xc, XW_C, N, B.cout + CO_C0))) return rc;
  if ((rc = gemm_dw(cb, B.dc2, DW_C, 64, B.xc1, XW_C1, N, B.cout + CO_C1))) return rc;
  if ((rc = gemm_dw(cb, B.dr, DW_R, 4, B.xc2, XW_C2, N, B.cout + CO_C2))) return rc;
  big_unpack_kernel<<<64, 256, 0, st>>>(B.cout, P, G);
  return check_cuda(cudaGetLastError(), "big_unpack_kernel");
}

}  // namespace fnr

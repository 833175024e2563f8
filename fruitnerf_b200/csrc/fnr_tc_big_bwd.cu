// Backward of the fruit_nerf_big family on the tensor cores, in two kernels of this library (no cuBLAS):
//   1. tc_big_backward_chain_kernel: per 128-point tile, recompute the activations from the stashed encoding (4 GEMM round
//      trips, bit-identical to fnr_tc_big.cu, ReLU masks kept as register bitmasks), then back-propagate dY through the
//      colour / semantic / base MLPs (dX = dY W with the forward weight tiles read MN-major, A operands in tensor memory),
//      scatter the encoding gradient into the hash-table gradient and the appearance-embedding gradient -- and write every
//      layer's input X and pre-activation gradient dY to scratch AS THE OPERAND TILES OF THE WEIGHT-GRADIENT GEMMS: bf16 hi / lo
//      splits (the same packed words the thread just wrote to tensor memory) in the canonical core-matrix layout
//      [8-feature chunk][64 points][16 B], grouped into four stage blocks per 64-point record (layout below).
//   2. tc_big_dw_kernel: the weight / bias gradients dW = dY^T X, reduction over all points.  Persistent, one CTA per SM: a
//      producer thread streams the stage blocks with TMA bulk copies (cp.async.bulk global -> shared, mbarrier complete_tx) into
//      a two-slot shared-memory ring, an MMA thread issues tcgen05.mma straight on the landed tiles (both operands MN-major,
//      K = points; hi*hi + lo*hi + hi*lo), and ALL accumulators (448 tensor-memory columns) stay resident for the whole kernel;
//      they are flushed once per CTA with atomics into the torch-layout gradient tensors.  big_fold_kernel finishes the folded
//      semantic tail (d W_sem2 = head_w (x) v, d head_w = W_sem2 v + b_sem2 s, ...).
// The 128-wide semantic layers make accumulators resident in the CHAIN kernel (fnr_tc_bwd.cu) impossible here: 448 accumulator
// columns plus the 384 columns of the chain exceed the 512 available.  Scratch: 4.3 KB per point.
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
constexpr int kChainLevelsDefault = 8;  // level-major mode: coarsest levels still scattered by the chain kernel's scatter warps (at most 8; measured
                                        // backward phase: 0 levels 2.125 ms, 4: 2.080, 6: 2.054, 8: 1.998; 10 / 12 / 16 levels: 2.15 / 2.17 / 2.32)
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

// ---- scratch: one record per 64 points; a blob of F features = [hi: F/8 chunks][lo: F/8 chunks], chunk = 64 points x 16 B.
// Four stage blocks per record, each the operand set of the GEMMs the dW kernel issues on it (chunk counts in brackets):
//   stage 1 (S1)        : DZ2 [16]  XZ1 [16]
//   stage 2 (F, S0)     : XZ2 [16]  DR [2]  DZ1 [16]  XG [4]
//   stage 3 (C0, C1)    : DC1 [8]   XC [6]  XG' [4]   DC2 [8]  XC1 [8]
//   stage 4 (C2, B0, B1): XC2 [8]   DR' [2] DH [8]    XE [4]   XH [8]  DOUT [4]
// X tiles: XE encoding, XH h, XG [h0 slot | geo 30 | 1] (the pad position of the forward carries a constant 1: the weight rows of
// that position are zero, and its column of the product is the bias gradient), XZ1 / XZ2 semantic hidden layers, XC [sh 16 | app 32],
// XC1 / XC2 colour hidden layers; dY tiles: DH, DOUT (base), DZ1, DZ2 (semantic), DC1, DC2 (colour), DR = [d rgb_pre 3, 0, d logit, 0..].
constexpr int kRecPoints = 64, kChunk = kRecPoints * 16;
constexpr int blob_bytes(int chunks) { return 2 * chunks * kChunk; }
constexpr int ST1 = 0, O_DZ2 = ST1, O_XZ1 = O_DZ2 + blob_bytes(16), ST1_BYTES = O_XZ1 + blob_bytes(16) - ST1;
constexpr int ST2 = ST1 + ST1_BYTES, O_XZ2 = ST2, O_DR = O_XZ2 + blob_bytes(16), O_DZ1 = O_DR + blob_bytes(2), O_XG = O_DZ1 + blob_bytes(16),
              ST2_BYTES = O_XG + blob_bytes(4) - ST2;
constexpr int ST3 = ST2 + ST2_BYTES, O_DC1 = ST3, O_XC = O_DC1 + blob_bytes(8), O_XG2 = O_XC + blob_bytes(6), O_DC2 = O_XG2 + blob_bytes(4),
              O_XC1 = O_DC2 + blob_bytes(8), ST3_BYTES = O_XC1 + blob_bytes(8) - ST3;
constexpr int ST4 = ST3 + ST3_BYTES, O_XC2 = ST4, O_DR2 = O_XC2 + blob_bytes(8), O_DH = O_DR2 + blob_bytes(2), O_XE = O_DH + blob_bytes(8),
              O_XH = O_XE + blob_bytes(4), O_DOUT = O_XH + blob_bytes(8), ST4_BYTES = O_DOUT + blob_bytes(4) - ST4;
constexpr int kRecBytes = ST4 + ST4_BYTES;
constexpr int kStageMax = ST2_BYTES;
static_assert(ST1_BYTES <= kStageMax && ST3_BYTES <= kStageMax && ST4_BYTES <= kStageMax, "ring slot size");
static_assert(kRecBytes % 128 == 0 && ST2 % 128 == 0 && ST3 % 128 == 0 && ST4 % 128 == 0, "bulk-copy alignment");
constexpr int kFoldFloats = 256;  // v[128] = sum dlogit z2, s = sum dlogit (+ padding): reduced over CTAs, consumed by big_fold_kernel

struct ChainArgs {
  KField F;
  KParams P;
  KParams G;
  KRays Rr;
  const float* point_grads;
  const float* stash;
  const float* sample_rgb;
  float* denc;       // [N,32] encoding gradient for the level-major scatter of tc_big_dw_kernel, or NULL: scatter warps of this kernel
  int chain_levels;  // with denc: the coarsest chain_levels levels are still scattered here, tile by tile, by the (otherwise idle) scatter
                     // warps -- their slices of the gradient table are small enough to stay L2-resident -- and tc_big_dw_kernel starts above
  int chain_merge;   // 0: full segmented scan on those levels, 1: one merge round
  uint8_t* records;  // scratch: ceil(N / 64) records of kRecBytes
  float* fold;       // [kFoldFloats] zero-initialised: the chain adds s = sum dlogit at [128]
  int debug_flags;   // FNR_DEBUG_BWD (timing experiments only): bit 0 skip the table scatter, bit 2 skip the X / dY stores
};

// 256-bit store (STG.E.ENL2.256): one whole 32-byte sector per lane and instruction; dst must be 32-byte aligned
__device__ __forceinline__ void store8(float* dst, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6),
               "f"(a7)
               : "memory");
}
// 16 values -> packed bf16 hi / lo words (two values per word)
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&h)[8], uint32_t (&l)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], h[q], l[q]);
}
// K elements [k0, k0 + 16) of this thread's row -> A tile (depth K) in tensor memory
template <int K>
__device__ __forceinline__ void st_a16p(uint32_t ab, int k0, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
  tmem_st8(ab + (k0 >> 1), h);
  tmem_st8(ab + K / 2 + (k0 >> 1), l);
}
template <int K>
__device__ __forceinline__ void st_a16(uint32_t ab, int k0, const float (&v)[16]) {
  uint32_t h[8], l[8];
  pack16(v, h, l);
  st_a16p<K>(ab, k0, h, l);
}
// features [k0, k0 + 16) of this thread's point -> blob (chunks = its feature count / 8) of the point's record; `rec` already
// points at the record + (point % 64) * 16.  A warp writes 512 contiguous bytes per instruction.
__device__ __forceinline__ void st_blob16(uint8_t* rec, int blob_off, int chunks, int k0, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
  uint8_t* p = rec + blob_off + (k0 >> 3) * kChunk;
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + kChunk) = make_uint4(h[4], h[5], h[6], h[7]);
  p += chunks * kChunk;
  *reinterpret_cast<uint4*>(p) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(p + kChunk) = make_uint4(l[4], l[5], l[6], l[7]);
}

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
    if (!a.denc || a.chain_levels > 0) {  // (with a.denc the encoding gradient goes to global memory and tc_big_dw_kernel scatters the levels
                                          // from chain_levels on, level by level)
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
      if (do_scatter && a.denc) {
#pragma unroll 1
        for (int l = 0; l < a.chain_levels; ++l) {
          float g0 = 0.f, g1 = 0.f;
#pragma unroll
          for (int ll = 0; ll < 8; ++ll)
            if (ll == l) {
              g0 = g[2 * ll];
              g1 = g[2 * ll + 1];
            }
          if (a.chain_merge) scatter_level_merged1(gtab, pos, live, g0, g1, l, F.scalings[l], F.log2T, hmask, lane);
          else scatter_level_aggregated(gtab, pos, live, g0, g1, l, F.scalings[l], F.log2T, hmask, lane);
        }
      } else if (do_scatter) {
#pragma unroll 1
        for (int l = 0; l < kAggLevels; ++l) {
          float g0 = 0.f, g1 = 0.f;
#pragma unroll
          for (int ll = 0; ll < kAggLevels; ++ll)
            if (ll == l) {
              g0 = g[2 * ll];
              g1 = g[2 * ll + 1];
            }
          scatter_level_aggregated(gtab, pos, live, g0, g1, l, F.scalings[l], F.log2T, hmask, lane);
        }
        if (live) {
#pragma unroll
          for (int l = kAggLevels; l < 16; ++l) scatter_level_direct(gtab, pos, g[2 * l], g[2 * l + 1], l, F.scalings[l], F.log2T, hmask);
        }
      }
    }
    }
  } else {
  // ================= compute warps =================
  reg_inc<216>();
  // narrow bias gradients as per-thread running sums: half 0 -> d rgb_pre (col2 bias) and d logit (folded tail), both halves -> dout
  float acc_dr[4] = {0.f, 0.f, 0.f, 0.f}, acc_dout[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc_dout[q] = 0.f;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long p = tile * 128 + row;
    const bool in_range = p < N;
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
    // this point's row in its 64-point record (rows past N carry dY = 0, so whatever X they hold contributes nothing)
    uint8_t* const rec = a.records + (size_t)(2 * tile + (row >> 6)) * kRecBytes + (row & 63) * 16;
    const bool valid = !(a.debug_flags & 4);
    uint32_t ph[8], pl[8];

    // ---- R1: encoding (stash) -> A0 ; X_enc ----
    {
      float enc[16];
      const float4* s4 = reinterpret_cast<const float4*>(a.stash + (size_t)pc * ENC + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 u = __ldg(s4 + q);
        enc[4 * q] = u.x; enc[4 * q + 1] = u.y; enc[4 * q + 2] = u.z; enc[4 * q + 3] = u.w;
      }
      pack16(enc, ph, pl);
      st_a16p<K_BASE0>(tr + R_A0, 16 * half, ph, pl);
      if (valid) st_blob16(rec, O_XE, 4, 16 * half, ph, pl);
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
      pack16(v, ph, pl);
      st_a16p<K_BASE1>(tr + R_A1, 32 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_XH, 8, 32 * half + 16 * j, ph, pl);
    }
    if (half == 0) {
      float sh[16];
      sh_degree4(__ldg(d), __ldg(d + 1), __ldg(d + 2), sh);
      pack16(sh, ph, pl);
      st_a16p<K_COL0>(tr + R_A0, 0, ph, pl);
      if (valid) st_blob16(rec, O_XC, 6, 0, ph, pl);
    } else {
      const float* app = (F.appearance_mode == FNR_APP_PER_CAMERA) ? P.app_embedding + (size_t)cam * APP : nullptr;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = app ? __ldg(app + 16 * j + q) : 0.f;
        pack16(v, ph, pl);
        st_a16p<K_COL0>(tr + R_A0, SHD + 16 * j, ph, pl);
        if (valid) st_blob16(rec, O_XC, 6, SHD + 16 * j, ph, pl);
      }
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
      pack16(g, ph, pl);
      st_a16p<K_SEM0>(tr + R_A1 + 64, 16 * half, ph, pl);
      st_a16p<K_COL0>(tr + R_A0, SHD + APP + 16 * half, ph, pl);
      if (valid) {
        if (half == 1) {  // the stored copy carries a constant 1 in the pad position (bias-gradient column of S0 / C0)
          g[15] = 1.0f;
          pack16(g, ph, pl);
        }
        st_blob16(rec, O_XG, 4, 16 * half, ph, pl);
        st_blob16(rec, O_XG2, 4, 16 * half, ph, pl);
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
      pack16(v, ph, pl);
      st_a16p<K_SEM1>(tr + R_A1, 64 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_XZ1, 16, 64 * half + 16 * j, ph, pl);
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
      pack16(v, ph, pl);
      st_a16p<K_COL1>(tr + R_A0, 32 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_XC1, 8, 32 * half + 16 * j, ph, pl);
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
      pack16(dz, ph, pl);
      st_a16p<K_SEM1>(tr + R_A1, 64 * half + 16 * j, ph, pl);
      if (valid) {
        st_blob16(rec, O_DZ2, 16, 64 * half + 16 * j, ph, pl);
        pack16(v, ph, pl);
        st_blob16(rec, O_XZ2, 16, 64 * half + 16 * j, ph, pl);
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
      if (valid) {
        pack16(v, ph, pl);
        st_blob16(rec, O_XC2, 8, 32 * half + 16 * j, ph, pl);
      }
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
      acc_dr[0] += dr[0];
      acc_dr[1] += dr[1];
      acc_dr[2] += dr[2];
      acc_dr[3] += d_logit;
      if (valid) {  // DR = [d rgb_pre 0..2, 0, d logit, 0 ...]: B operand of C2 (columns 0..2) and of the folded tail F (column 4)
        dr[4] = d_logit;
        pack16(dr, ph, pl);
        st_blob16(rec, O_DR, 2, 0, ph, pl);
        st_blob16(rec, O_DR2, 2, 0, ph, pl);
      }
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
      if (valid) {
        pack16(v, ph, pl);
        st_blob16(rec, O_DZ1, 16, 64 * half + 16 * j, ph, pl);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r[16];
      tmem_ld16(tr + R_D1 + 32 * half + 16 * j, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = ((mask_c2 >> (16 * j + q)) & 1u) ? __uint_as_float(r[q]) : 0.f;
      pack16(v, ph, pl);
      st_a16p<K_COL2>(tr + R_A0, 32 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_DC2, 8, 32 * half + 16 * j, ph, pl);
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
      pack16(v, ph, pl);
      st_a16p<K_COL1>(tr + R_A0, 32 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_DC1, 8, 32 * half + 16 * j, ph, pl);
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
      pack16(dout, ph, pl);
      st_a16p<N_BASE1>(tr + R_A1, 16 * half, ph, pl);
#pragma unroll
      for (int q = 0; q < 16; ++q) acc_dout[q] += dout[q];  // base1 bias gradient (its X tile h has no spare column)
      if (valid) st_blob16(rec, O_DOUT, 4, 16 * half, ph, pl);
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
      pack16(v, ph, pl);
      st_a16p<N_BASE0>(tr + R_A0, 32 * half + 16 * j, ph, pl);
      if (valid) st_blob16(rec, O_DH, 8, 32 * half + 16 * j, ph, pl);
    }
    FNR_ISSUE(issue_dx<N_BASE0, K_BASE0>(tb + R_D0, tb + R_A0, wBase + OFF_W_BASE0))

    // ---- epi10: d enc -> hand-off to the scatter warps ----
    FNR_WAIT()
    {
      uint32_t r[16];
      tmem_ld16(tr + R_D0 + 16 * half, r);
      tmem_ld_wait();
      if (a.denc) {
        if (in_range) {
          float* dst = a.denc + (size_t)p * ENC + 16 * half;  // 64-byte aligned: two whole sectors per thread
          store8(dst, __uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]), __uint_as_float(r[4]),
                 __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
          store8(dst + 8, __uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]), __uint_as_float(r[12]),
                 __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
        }
      }
      if (!a.denc || a.chain_levels > 0) {
        named_bar_sync(BAR_EMPTY, kCtaThreads);  // the scatter warps have copied the previous tile out
        float4* dst = reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 16 * half);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        if (half == 0) reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 32)[0] = make_float4(pos.x, pos.y, pos.z, in_range ? 1.f : 0.f);
      }
    }
    if (!a.denc || a.chain_levels > 0) named_bar_arrive(BAR_FULL, kCtaThreads);
    fence_before_sync();
  }
  // flush the running sums: warp reduce, one atomic per warp and value
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float sum = warp_sum_f(acc_dout[q]);
    const int n = 16 * half + q;
    if (lane == 0 && n < 1 + GEO && sum != 0.f) atomicAdd(G.base_b[1] + n, sum);
  }
  if (half == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sum = warp_sum_f(acc_dr[c]);
      if (lane == 0 && sum != 0.f) atomicAdd(c < 3 ? G.col_b[2] + c : a.fold + 128, sum);
    }
  }
  }  // compute warps
#undef FNR_ISSUE
#undef FNR_WAIT

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

// ======================================================================================================================
// tc_big_dw_kernel: dW = dY^T X over all points, operands streamed by TMA, accumulators resident in tensor memory
// ======================================================================================================================
constexpr int kDwThreads = 512;   // warp 0: TMA producer, warp 1: MMA issue, warps 2..15: level-major table scatter, warps 0..7: final flush
constexpr int kDwSlots = 2;
#ifndef FNR_DW_AGG_LEVELS
#define FNR_DW_AGG_LEVELS 9
#endif
constexpr int kDwAggLevels = FNR_DW_AGG_LEVELS;  // levels whose reds are run-length merged across the warp in the level-major scatter: the reds are bound by
                                                 // the chip-wide L2 atomic request rate (~65-80 G requests/s measured), the SMs of this kernel are otherwise idle
constexpr int DW_OFF_ONES = 0;                         // [2 chunks][64 points][16 B]: feature 0 = 1 (bias-gradient operand)
constexpr int DW_OFF_RING = 2 * kChunk;
constexpr int kDwSmem = DW_OFF_RING + kDwSlots * kStageMax + 1024;
static_assert(kDwSmem <= 227 * 1024, "shared-memory budget");
// accumulator columns (M = 128: lane = row; M = 64: row i in lane (i % 16) + 32 (i / 16))
constexpr int A_S1 = 0;      // [dz2 128 x z1 128]
constexpr int A_S1B = 128;   // [dz2 128 x 16]   column 0 = bias
constexpr int A_F = 144;     // [z2 128 x 16]    column 4 = v
constexpr int A_S0 = 160;    // [dz1 128 x 32]   columns 1..30 = geo, 31 = bias
constexpr int A_C0A = 192;   // [dc1 64 x 48]    sh | app
constexpr int A_C0B = 240;   // [dc1 64 x 32]    columns 1..30 = geo, 31 = bias
constexpr int A_C1 = 272;    // [dc2 64 x 64]
constexpr int A_C1B = 336;   // [dc2 64 x 16]    column 0 = bias
constexpr int A_C2 = 352;    // [c2 64 x 16]     columns 0..2 = colour outputs
constexpr int A_B0 = 368;    // [dh 64 x 32]
constexpr int A_B0B = 400;   // [dh 64 x 16]     column 0 = bias
constexpr int A_B1 = 416;    // [h 64 x 32]      columns 0..30 = base outputs
static_assert(A_B1 + 32 <= 512, "tensor-memory budget");

struct DwArgs {
  const uint8_t* records;
  long long num_records;
  KParams G;
  float* fold;  // [kFoldFloats]: v[128] (+= over CTAs); [128] = s comes from the chain kernel
  // level-major hash-table gradient scatter by the six otherwise idle warps (denc == NULL: the chain kernel scattered already)
  const float* denc;  // [N,32]
  int first_level;    // levels below were scattered by the chain kernel
  KField F;
  KRays Rr;
  long long num_points;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ACC[M, NB] (+)= A[64 points, M]^T B[64 points, NB]: both tiles MN-major (chunk stride 1024 B), K = 64 points = 4 K-steps.
// a / b: shared addresses of the hi blobs; a_lo / b_lo: byte offsets of the lo halves (0 = operand has no lo part).
template <int M, int NB>
__device__ __forceinline__ void issue_dw(uint32_t d_tmem, uint32_t a, uint32_t a_lo, uint32_t b, uint32_t b_lo, bool accumulate) {
  constexpr uint32_t idesc = idesc_bf16_f32(M, NB) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
#pragma unroll
  for (int ks = 0; ks < kRecPoints / 16; ++ks) {
    const uint64_t ah = smem_desc(a + ks * 256, 128, kChunk);
    const uint64_t bh = smem_desc(b + ks * 256, 128, kChunk);
    mma_ss(d_tmem, ah, bh, idesc, accumulate || ks > 0);
    if (a_lo) mma_ss(d_tmem, smem_desc(a + a_lo + ks * 256, 128, kChunk), bh, idesc, true);
    if (b_lo) mma_ss(d_tmem, ah, smem_desc(b + b_lo + ks * 256, 128, kChunk), idesc, true);
  }
}

__global__ void __launch_bounds__(kDwThreads, 1) tc_big_dw_kernel(const __grid_constant__ DwArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_full[kDwSlots], s_empty[kDwSlots], s_done;
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const KParams& G = a.G;

  if (warp == 0) tmem_alloc(&s_tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < kDwSlots; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 1);
    }
    mbar_init(&s_done, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 2 * kRecPoints; i += kDwThreads)  // ONES: chunk 0 = (1, 0, 0, ...) per point, chunk 1 = 0
    if (i < 2 * kRecPoints)
    *reinterpret_cast<uint4*>(smem + DW_OFF_ONES + i * 16) = make_uint4(i < kRecPoints ? 0x00003F80u : 0u, 0u, 0u, 0u);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = s_tmem_base;
  const uint32_t sb = smem_u32(smem);
  const long long my_records = a.num_records > blockIdx.x ? (a.num_records - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  constexpr int kStageOff[4] = {ST1, ST2, ST3, ST4};
  constexpr int kStageBytes[4] = {ST1_BYTES, ST2_BYTES, ST3_BYTES, ST4_BYTES};

  if (warp == 0) {
    // ---- TMA producer: one bulk copy per stage block into the ring
    if (elect_one_sync()) {
      long long it = 0;
      for (long long i = 0; i < my_records; ++i) {
        const uint8_t* rec = a.records + (size_t)(blockIdx.x + i * gridDim.x) * kRecBytes;
#pragma unroll
        for (int st = 0; st < 4; ++st, ++it) {
          const int slot = (int)(it % kDwSlots);
          const uint32_t use = (uint32_t)(it / kDwSlots);
          mbar_wait(&s_empty[slot], (use & 1u) ^ 1u);  // the MMAs that read this slot's previous contents have completed
          const uint32_t bar = smem_u32(&s_full[slot]);
          mbar_expect_tx(bar, (uint32_t)kStageBytes[st]);
          bulk_g2s(sb + DW_OFF_RING + slot * kStageMax, rec + kStageOff[st], (uint32_t)kStageBytes[st], bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issue: every stage block is the operand set of the GEMMs issued on it
    if (elect_one_sync()) {
      long long it = 0;
      const uint32_t ones = sb + DW_OFF_ONES;
      for (long long i = 0; i < my_records; ++i) {
        const bool acc = i > 0;
#pragma unroll
        for (int st = 0; st < 4; ++st, ++it) {
          const int slot = (int)(it % kDwSlots);
          const uint32_t use = (uint32_t)(it / kDwSlots);
          mbar_wait(&s_full[slot], use & 1u);
          fence_after_sync();
          const uint32_t base = sb + DW_OFF_RING + slot * kStageMax - kStageOff[st];  // record-relative offsets below
          if (st == 0) {
            issue_dw<128, 128>(tb + A_S1, base + O_DZ2, 16 * kChunk, base + O_XZ1, 16 * kChunk, acc);
            issue_dw<128, 16>(tb + A_S1B, base + O_DZ2, 16 * kChunk, ones, 0, acc);
          } else if (st == 1) {
            issue_dw<128, 16>(tb + A_F, base + O_XZ2, 16 * kChunk, base + O_DR, 2 * kChunk, acc);
            issue_dw<128, 32>(tb + A_S0, base + O_DZ1, 16 * kChunk, base + O_XG, 4 * kChunk, acc);
          } else if (st == 2) {
            issue_dw<64, 48>(tb + A_C0A, base + O_DC1, 8 * kChunk, base + O_XC, 6 * kChunk, acc);
            issue_dw<64, 32>(tb + A_C0B, base + O_DC1, 8 * kChunk, base + O_XG2, 4 * kChunk, acc);
            issue_dw<64, 64>(tb + A_C1, base + O_DC2, 8 * kChunk, base + O_XC1, 8 * kChunk, acc);
            issue_dw<64, 16>(tb + A_C1B, base + O_DC2, 8 * kChunk, ones, 0, acc);
          } else {
            issue_dw<64, 16>(tb + A_C2, base + O_XC2, 8 * kChunk, base + O_DR2, 2 * kChunk, acc);
            issue_dw<64, 32>(tb + A_B0, base + O_DH, 8 * kChunk, base + O_XE, 4 * kChunk, acc);
            issue_dw<64, 16>(tb + A_B0B, base + O_DH, 8 * kChunk, ones, 0, acc);
            issue_dw<64, 32>(tb + A_B1, base + O_XH, 8 * kChunk, base + O_DOUT, 4 * kChunk, acc);
          }
          mma_commit(&s_empty[slot]);  // arrives when the MMAs above have finished reading the slot
        }
      }
      mma_commit(&s_done);
    }
    __syncwarp();
  }

  else if (a.denc) {
    // ---- table scatter, LEVEL-MAJOR: all CTAs sweep level l together, so the reds of a sweep fall into ONE level's slice of the
    // gradient table (2^T rows x 8 B = 16.8 MB at T = 21), which stays L2-resident -- scattered from inside the chain kernel, tile by
    // tile over all 16 levels, they miss L2 (268 MB table) and every red is a DRAM read-modify-write.  This kernel's own work (a
    // TMA-fed HBM stream + tensor-core MMAs issued by two threads) leaves the SM's LSU / atomic path idle: the two overlap.
    const int sw = warp - 2;
    constexpr int kSW = kDwThreads / 32 - 2;
    const KField& F = a.F;
    const uint32_t hmask = (1u << F.log2T) - 1u;
    float2* const gtab = reinterpret_cast<float2*>(G.hash_table);
    const long long N = a.num_points;
    const long long chunks = (N + 31) / 32;
    const int S = a.Rr.S;
#pragma unroll 1
    for (int l = a.first_level; l < F.L; ++l) {
      const float scale = F.scalings[l];
#pragma unroll 1
      for (long long c = (long long)blockIdx.x * kSW + sw; c < chunks; c += (long long)gridDim.x * kSW) {
        const long long p = c * 32 + lane;
        const bool live = p < N;
        const long long pc = live ? p : N - 1;
        const int ray = (int)(pc / S);
        bool sel;
        const Vec3 pos = field_position(a.Rr.origins + 3 * (size_t)ray, a.Rr.directions + 3 * (size_t)ray, __ldg(a.Rr.starts + pc),
                                        __ldg(a.Rr.ends + pc), F.position_mode, F.aabb, sel);
        const float2 g = __ldg(reinterpret_cast<const float2*>(a.denc + (size_t)pc * ENC) + l);
        // full segmented scan here (the 14 scatter warps of this kernel hide its shuffles; the one-round merge the small backward
        // uses measured 2.34 ms against 2.13 ms for the backward phase, no merge 2.52 ms -- tools/r2/run18.sh)
        if (l < kDwAggLevels) scatter_level_aggregated(gtab, pos, live, g.x, g.y, l, scale, F.log2T, hmask, lane);
        else if (live) scatter_level_direct(gtab, pos, g.x, g.y, l, scale, F.log2T, hmask);
      }
    }
  }

  // ---- flush: every accumulator once per CTA, atomics into the torch-layout gradient tensors
  if (my_records > 0 && warp < 8) {
    mbar_wait(&s_done, 0);
    fence_after_sync();
    const int quarter = warp & 3, half = warp >> 2;  // 8 warps: lane quarter x column half
    const uint32_t tr = tb + ((uint32_t)(quarter * 32) << 16);
    const int row128 = quarter * 32 + lane;            // M = 128 accumulators: row = lane
    const bool own64 = lane < 16;                      // M = 64 accumulators: rows live in lanes 0..15 of every quarter
    const int row64 = 16 * quarter + lane;
    uint32_t r[16];
    // S1: d W_sem1[n][k], n = row, k = column
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int c0 = 64 * half + 16 * j;
      tmem_ld16(tr + A_S1 + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) atomicAdd(G.sem_w[1] + row128 * SW + c0 + q, __uint_as_float(r[q]));
    }
    tmem_ld16(tr + (half == 0 ? A_S1B : A_F), r);
    tmem_ld_wait();
    if (half == 0) atomicAdd(G.sem_b[1] + row128, __uint_as_float(r[0]));
    else atomicAdd(a.fold + row128, __uint_as_float(r[4]));  // v[k] = sum dlogit z2[k]
    // S0: d W_sem0[n][geo], bias in column 31
    tmem_ld16(tr + A_S0 + 16 * half, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int kk = 16 * half + q;
      if (kk >= 1 && kk <= GEO) atomicAdd(G.sem_w[0] + row128 * GEO + (kk - 1), __uint_as_float(r[q]));
      else if (kk == 31) atomicAdd(G.sem_b[0] + row128, __uint_as_float(r[q]));
    }
    // M = 64 accumulators
    float* gw0 = G.col_w[0] + row64 * CIN;
#pragma unroll 1
    for (int j = half; j < 3; j += 2) {  // C0A: 48 columns [sh 16 | app 32] -> torch order [sh | geo | app]
      tmem_ld16(tr + A_C0A + 16 * j, r);
      tmem_ld_wait();
      if (own64) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int k = 16 * j + q;
          atomicAdd(gw0 + (k < SHD ? k : SHD + GEO + (k - SHD)), __uint_as_float(r[q]));
        }
      }
    }
    tmem_ld16(tr + A_C0B + 16 * half, r);
    tmem_ld_wait();
    if (own64) {
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int kk = 16 * half + q;
        if (kk >= 1 && kk <= GEO) atomicAdd(gw0 + SHD + (kk - 1), __uint_as_float(r[q]));
        else if (kk == 31) atomicAdd(G.col_b[0] + row64, __uint_as_float(r[q]));
      }
    }
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {  // C1
      const int c0 = 32 * half + 16 * j;
      tmem_ld16(tr + A_C1 + c0, r);
      tmem_ld_wait();
      if (own64) {
#pragma unroll
        for (int q = 0; q < 16; ++q) atomicAdd(G.col_w[1] + row64 * H + c0 + q, __uint_as_float(r[q]));
      }
    }
    tmem_ld16(tr + (half == 0 ? A_C1B : A_C2), r);
    tmem_ld_wait();
    if (own64) {
      if (half == 0) atomicAdd(G.col_b[1] + row64, __uint_as_float(r[0]));
      else {
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(G.col_w[2] + c * H + row64, __uint_as_float(r[c]));  // rows = c2 features
      }
    }
    tmem_ld16(tr + A_B0 + 16 * half, r);
    tmem_ld_wait();
    if (own64) {
#pragma unroll
      for (int q = 0; q < 16; ++q) atomicAdd(G.base_w[0] + row64 * ENC + 16 * half + q, __uint_as_float(r[q]));
    }
    tmem_ld16(tr + (half == 0 ? A_B0B : A_B1), r);
    tmem_ld_wait();
    if (own64) {
      if (half == 0) atomicAdd(G.base_b[0] + row64, __uint_as_float(r[0]));
      else {
#pragma unroll
        for (int q = 0; q < 16; ++q) atomicAdd(G.base_w[1] + q * H + row64, __uint_as_float(r[q]));  // rows = h features, columns = outputs 0..15
      }
    }
    if (half == 1) {
      tmem_ld16(tr + A_B1 + 16, r);
      tmem_ld_wait();
      if (own64) {
#pragma unroll
        for (int q = 0; q < 15; ++q) atomicAdd(G.base_w[1] + (16 + q) * H + row64, __uint_as_float(r[q]));  // outputs 16..30 (31 is the pad)
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

// folded semantic tail from v[k] = sum dlogit z2[k] and s = sum dlogit:
//   logit = head_w . (W_sem2 z2 + b_sem2) + head_b
__global__ void __launch_bounds__(256) big_fold_kernel(const float* __restrict__ fold, KParams P, KParams G) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const float* vv = fold;
  const float s = fold[128];
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

}  // namespace

size_t tc_big_backward_scratch_bytes(long long num_points) {
  const long long records = 2 * ((num_points + 127) / 128);
  return (size_t)records * kRecBytes + (size_t)num_points * ENC * sizeof(float) + (size_t)kFoldFloats * sizeof(float) + 8192;
}

bool tc_big_backward_supported(const KField& F, const KFieldBwd& B) {
  return B.stash_encoding != nullptr && B.sample_rgb != nullptr && B.extra != nullptr && !F.pass_semantic_gradients &&
         (F.appearance_mode == FNR_APP_PER_CAMERA || F.appearance_mode == FNR_APP_ZEROS);
}

int launch_tc_big_field_backward(const KField& F, const KParams& P, const KParams& G, const KRays& Rr, const KFieldBwd& Bw, cudaStream_t st) {
  const long long N = (long long)Rr.R * Rr.S;
  if (N == 0) return FNR_OK;
  if (Bw.extra_bytes < tc_big_backward_scratch_bytes(N)) {
    set_error("scratch too small for the big-family tensor-core backward");
    return FNR_ERR_INVALID_ARGUMENT;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_big_backward_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_big_backward_chain_kernel)");
    e = cudaFuncSetAttribute(tc_big_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_big_dw_kernel)");
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
  const long long tiles = (N + 127) / 128;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(Bw.extra) + 255) & ~(uintptr_t)255);
  a.fold = reinterpret_cast<float*>(base);
  a.records = base + 1024;
  // FNR_BIG_BWD_MODE=0: round-1 placement of the table scatter (scatter warps inside the chain kernel), for A/B timing
  static const bool level_major = !(getenv("FNR_BIG_BWD_MODE") && atoi(getenv("FNR_BIG_BWD_MODE")) == 0);
  a.denc = level_major ? reinterpret_cast<float*>(base + 1024 + (size_t)(2 * tiles) * kRecBytes) : nullptr;
  // FNR_BIG_CHAIN_LEVELS / FNR_BIG_CHAIN_MERGE: A/B switches of the hybrid placement (defaults = the measured best)
  static const int chain_levels = getenv("FNR_BIG_CHAIN_LEVELS") ? atoi(getenv("FNR_BIG_CHAIN_LEVELS")) : kChainLevelsDefault;
  static const int chain_merge = getenv("FNR_BIG_CHAIN_MERGE") ? atoi(getenv("FNR_BIG_CHAIN_MERGE")) : 0;
  a.chain_levels = level_major ? (chain_levels < 0 ? 0 : chain_levels > 8 ? 8 : chain_levels) : 0;
  a.chain_merge = chain_merge;
  if (int rc = check_cuda(cudaMemsetAsync(a.fold, 0, kFoldFloats * sizeof(float), st), "cudaMemsetAsync(fold)")) return rc;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc_big_backward_chain_kernel<<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  if (int rc = check_launch("tc_big_backward_chain_kernel")) return rc;
  if (a.debug_flags & 4) return FNR_OK;  // timing experiment: no X / dY tiles were written

  DwArgs d;
  d.records = a.records;
  d.num_records = 2 * tiles;
  d.G = G;
  d.fold = a.fold;
  d.denc = (a.debug_flags & 1) ? nullptr : a.denc;
  d.first_level = a.chain_levels;
  d.F = F;
  d.Rr = Rr;
  d.num_points = N;
  const int dgrid = (int)(d.num_records < sm_count() ? d.num_records : sm_count());
  tc_big_dw_kernel<<<dgrid, kDwThreads, kDwSmem, st>>>(d);
  if (int rc = check_launch("tc_big_dw_kernel")) return rc;
  big_fold_kernel<<<32, 256, 0, st>>>(a.fold, P, G);
  return check_launch("big_fold_kernel");
}

}  // namespace fnr

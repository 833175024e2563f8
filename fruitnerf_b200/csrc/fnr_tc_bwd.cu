// Field backward on the 5th-gen tensor cores (sm_100a): for each tile of 128 sample points
//   recompute the activations (5 GEMM round trips, as the fused forward),
//   back-propagate through the colour / semantic / base MLPs   dX = dY W        (W read MN-major from the
//                                                                              same tile the forward uses),
//   accumulate every weight gradient                           dW = dY^T X      (both activation tiles read
//                                                                              MN-major; M=64 accumulators
//                                                                              RESIDENT in TMEM for the whole
//                                                                              kernel, flushed once per CTA),
//   scatter the encoding gradient into the hash-table gradient (red.global.add.v2.f32).
// Autograd semantics of fruit_nerf/fruit_field.py:168-281 (semantic branch sees detach(geo), trunc_exp
// backward clamps the exponent to [-15, 15]); upstream per-sample gradients come from the compositing
// backward (simt_composite_backward_kernel).
//
// One persistent CTA of 256 threads per SM: the thread pair (r, r+128) owns point r of the tile (TMEM lane r);
// the two threads split every 64-column epilogue and the 16 hash levels in halves (two warps may touch the
// same TMEM lane quarter), which doubles the loads / TMEM reads in flight of the single MMA slot.  All operands
// are bf16 hi/lo splits (3 MMAs per K-step, ~2^-16 per product); the K=16 "[geo | 1]" operand of the
// semantic branch is chunks 6-7 of the colour-input tile, so no activation is stored twice.
// Bias gradients come for free from constant-1 columns (padding columns of the geo / colour-input tiles,
// a 128x16 ONES tile for the K-exact layers) or from per-thread running sums (3- and 16-wide layers).
#include <cstdio>
#include <cstdlib>
#include "fnr_common.cuh"
#include "fnr_kernels.h"
#include "fnr_tc_common.cuh"

namespace fnr {
using namespace tc;
using namespace tcx;

namespace {

constexpr int kTC = 256;  // compute threads: 2 threads per point (column halves), 8 warps
#ifndef FNR_BWD_SCATTER_WARPS
#define FNR_BWD_SCATTER_WARPS 4
#endif
constexpr int kScatterWarps = FNR_BWD_SCATTER_WARPS;  // 4: one thread per point, all 16 levels; 8: two threads per point, 8 levels each
constexpr int kTS = 32 * kScatterWarps;  // scatter threads: warps that only issue the hash-table gradient reds
constexpr int kScatterRegs = 72;  // setmaxnreg budgets: kTS * kScatterRegs + kTC * kComputeRegs <= 64 K registers
constexpr int kComputeRegs = kScatterWarps == 4 ? 216 : 184;
static_assert(kScatterWarps == 4 || kScatterWarps == 8, "scatter warps");
constexpr int kT = kTC + kTS;
static_assert(kTS * kScatterRegs + kTC * kComputeRegs <= 65536 && kScatterRegs % 8 == 0 && kComputeRegs % 8 == 0, "register budget");
constexpr int GEO = 15, ENC = 32, H = 64, APP = 32, SHD = 16;

// ---- shared-memory map (bytes) ------------------------------------------------------------------
constexpr int OFF_W0 = 0;                                  // [64 rows][32]  hi,lo
constexpr int OFF_W1 = OFF_W0 + 2 * wbytes(64, 32);        // [16][64]
constexpr int OFF_WS0 = OFF_W1 + 2 * wbytes(16, 64);       // [64][16]
constexpr int OFF_WC0 = OFF_WS0 + 2 * wbytes(64, 16);      // [64][64], K order [sh|app|geo|0]
constexpr int OFF_WC1 = OFF_WC0 + 2 * wbytes(64, 64);      // [64][64]
constexpr int OFF_WC2 = OFF_WC1 + 2 * wbytes(64, 64);      // [16][64]
constexpr int OFF_F32 = OFF_WC2 + 2 * wbytes(16, 64);
// fp32 block: b0[64] b1[16] bs0[64] bc0[64] bc1[64] fold[64] reduce scratch[64]
constexpr int F_B0 = 0, F_B1 = 64, F_BS0 = 80, F_BC0 = 144, F_BC1 = 208, F_FOLD = 272, F_RED = 336, F_COUNT = 416;  // F_RED: v[64] + 4 per-warp partial sums
constexpr int TILE64 = 2 * 128 * 64 * 2;                   // 32 KB: K=64 tile, hi then lo
constexpr int OFF_H = OFF_F32 + F_COUNT * 4;               // h1 (hi | lo)
constexpr int OFF_CIN = OFF_H + TILE64;                    // colour input [sh | app | geo | 1]; chunks 6,7 double as the
                                                           // K=16 "[geo | 1]" tile of the semantic branch
constexpr int OFF_A = OFF_CIN + TILE64;                    // enc (K=32) -> z1 -> c2 -> enc again
constexpr int OFF_C1 = OFF_A + TILE64;
constexpr int OFF_DY = OFF_C1 + TILE64;
constexpr int OFF_ONES = OFF_DY + TILE64;                  // [128][16] bf16, column 0 = 1
constexpr int OFF_D16 = OFF_ONES + 128 * 16 * 2;           // 16-wide dY tile (hi 4 KB, lo 4 KB)
constexpr int OFF_END = OFF_D16 + 2 * 128 * 16 * 2;
constexpr int kSmem = OFF_END;
// denc hand-off to the scatter warps lives in the C1 tile, which is idle between a tile's T6 and the next tile's T3
constexpr int STAGE_STRIDE = 36;  // floats per row (32 denc + pos xyz + valid), padded: conflict-free 16-byte accesses
constexpr int GEO_CHUNK = 6 * 2048;                        // byte offset of chunk 6 inside a K=64 tile half
static_assert(kSmem + 64 <= 227 * 1024, "shared-memory budget");
static_assert(OFF_H % 16 == 0 && OFF_D16 % 16 == 0, "tile alignment");

// ---- tensor-memory map (columns) ----------------------------------------------------------------
constexpr int C_R0 = 0, C_R1 = 64, C_R2 = 80;              // work regions (64, 16, 64)
constexpr int C_AV = 144;    // z1^T dlogit            [64 x 16]
constexpr int C_AS0 = 160;   // dz1^T [geo|1]          [64 x 16]
constexpr int C_AC2 = 176;   // c2^T do3               [64 x 16]
constexpr int C_AC1 = 192;   // dc2^T c1               [64 x 64]
constexpr int C_AC1B = 256;  // dc2^T ones             [64 x 16]
constexpr int C_AC0 = 272;   // dc1^T cin              [64 x 64]
constexpr int C_AB1 = 336;   // h1^T dout16            [64 x 16]
constexpr int C_AB0 = 352;   // dh1^T enc              [64 x 32]
constexpr int C_AB0B = 384;  // dh1^T ones             [64 x 16]

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
  return idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// dX[128, NOUT] = dY[128, KR] * W[KR rows, NOUT cols]   (A K-major, B MN-major view of the forward tile)
template <int KR, int NOUT>
__device__ __forceinline__ void issue_dx(uint32_t d_tmem, uint32_t a_hi, uint32_t w_hi) {
  constexpr uint32_t idesc = idesc_mn(128, NOUT, 0, 1);
  constexpr uint32_t a_lo = 128 * KR * 2, w_lo = KR * NOUT * 2;
#pragma unroll
  for (int ks = 0; ks < KR / 16; ++ks) {
    const uint64_t ah = smem_desc(a_hi + ks * 4096, 2048, 128);
    const uint64_t al = smem_desc(a_hi + a_lo + ks * 4096, 2048, 128);
    const uint64_t wh = smem_desc(w_hi + ks * 256, 128, KR * 16);
    const uint64_t wl = smem_desc(w_hi + w_lo + ks * 256, 128, KR * 16);
    mma_ss(d_tmem, ah, wh, idesc, ks > 0);
    mma_ss(d_tmem, al, wh, idesc, true);
    mma_ss(d_tmem, ah, wl, idesc, true);
  }
}

// ACC[FA=64, FB] (+)= A[128, 64]^T * B[128, FB]   (both row-per-point tiles read MN-major, K = 128 points).
// a_lo / b_lo: shared addresses of the lo halves (0 = operand not split).
__device__ int g_skip_dw;
template <int FB>
__device__ __forceinline__ void issue_dw(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, bool accumulate) {
  constexpr uint32_t idesc = idesc_mn(64, FB, 1, 1);
  if (g_skip_dw) return;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t ah = smem_desc(a_hi + ks * 256, 128, 2048);
    const uint64_t bh = smem_desc(b_hi + ks * 256, 128, 2048);
    mma_ss(d_tmem, ah, bh, idesc, accumulate || ks > 0);
    if (a_lo) mma_ss(d_tmem, smem_desc(a_lo + ks * 256, 128, 2048), bh, idesc, true);
    if (b_lo) mma_ss(d_tmem, ah, smem_desc(b_lo + ks * 256, 128, 2048), idesc, true);
  }
}

// ---- column-split epilogues: the two threads (row, half=0/1) of a row each own 32 of the 64 columns ----
// v = f(column n, accumulator value); stores chunks 4*half .. 4*half+3 of a K=64 hi/lo tile.
template <class Fn>
__device__ __forceinline__ void epi32(uint32_t taddr32, uint8_t* tile, int row, int half, Fn f) {
  uint32_t r[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) tmem_ld8(taddr32 + 8 * j, r[j]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = f(32 * half + 8 * j + q, __uint_as_float(r[j][q]));
    store_chunk(tile, 128 * 64 * 2, row, 4 * half + j, c);
  }
}

// 8-bit ReLU mask of chunk j of a stored K=64 tile row (hi half): bit q set iff activation 8j+q > 0.
__device__ __forceinline__ uint32_t relu_mask8(const uint8_t* tile_hi, int row, int j) {
  const uint4 u = *reinterpret_cast<const uint4*>(tile_hi + j * 2048 + row * 16);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (w[q] & 0x7FFFu) m |= 1u << (2 * q);
    if (w[q] & 0x7FFF0000u) m |= 1u << (2 * q + 1);
  }
  return m;
}

// dX epilogue: v = acc * relu'(stored activation of the same column) -> DY tile.
__device__ __forceinline__ void epi32_masked(uint32_t taddr32, const uint8_t* act_tile, uint8_t* dst, int row, int half) {
  uint32_t r[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) tmem_ld8(taddr32 + 8 * j, r[j]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t m = relu_mask8(act_tile, row, 4 * half + j);
    float c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = ((m >> q) & 1u) ? __uint_as_float(r[j][q]) : 0.f;
    store_chunk(dst, 128 * 64 * 2, row, 4 * half + j, c);
  }
}

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int BAR_COMPUTE = 1, BAR_FULL = 2, BAR_EMPTY = 3;
#ifndef FNR_MERGE_END
#define FNR_MERGE_END 8
#endif
// Levels 0 .. kMergeEnd-1 of the table-gradient scatter merge neighbouring lanes of a same-cell run before issuing (see the scatter
// warps below); measured on the bench batch (tools/r2/run16.sh, run17.sh): no merge on the coarse levels 0.91 ms, the round-1 full
// segmented scan (5 shuffle rounds) on levels 0-3 0.843 ms, one round on levels 0-5 / 0-7 / 0-15 0.814 ms.
constexpr int kMergeEnd = FNR_MERGE_END;
#ifndef FNR_BWD_OFFLOAD
#define FNR_BWD_OFFLOAD 6
#endif
// The finest kOffload levels of a tile's table gradient are scattered by the chain warps 4-7 themselves (they hold those columns of
// the encoding gradient in registers at T9), one level inside each of the long MMA waits of the NEXT tile (T4, T6, T7, T9; + T5, T8; + T2, T3):
// those warps idle there, while the four scatter warps are the critical path of the kernel.
constexpr int kOffload = FNR_BWD_OFFLOAD;
static_assert(kOffload == 0 || kOffload == 4 || kOffload == 6 || kOffload == 8, "offloaded levels");
static_assert(kOffload == 0 || kScatterWarps == 4, "level offload assumes one scatter thread per point");

struct BwdArgs {
  KField F;
  KParams P;
  KParams G;
  KRays Rr;
  const float* point_grads;     // [N,5]  d_density, d_rgb[3], d_logit
  const float* stash_encoding;  // [N,32]
  const float* sample_rgb;      // [N,3]
  int debug_flags;              // FNR_DEBUG_BWD env (bit 0: skip table scatter, bit 1: skip dW GEMMs) -- timing experiments only
};

__global__ void __launch_bounds__(kT, 1) tc_field_backward_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_compute = tid < kTC;
  const int row = tid & 127;          // point (TMEM lane) of this thread (compute) / point it scatters (scatter warps)
  const int half = (tid >> 7) & 1;    // compute threads: which 32 of 64 accumulator columns this thread owns
  const KParams& P = a.P;
  const KParams& G = a.G;
  const KField& F = a.F;
  float* sf = reinterpret_cast<float*>(smem + OFF_F32);

  if (warp == 0) tmem_alloc(&s_tmem_base, 512);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  stage_weight<kT, 64, 32>(smem + OFF_W0, [&](int n, int k) { return __ldg(P.base_w[0] + n * ENC + k); });
  stage_weight<kT, 16, 64>(smem + OFF_W1, [&](int n, int k) { return __ldg(P.base_w[1] + n * H + k); });
  stage_weight<kT, 64, 16>(smem + OFF_WS0, [&](int n, int k) { return k < GEO ? __ldg(P.sem_w[0] + n * GEO + k) : 0.f; });
  stage_weight<kT, 64, 64>(smem + OFF_WC0, [&](int n, int k) {
    const float* w = P.col_w[0] + n * (SHD + GEO + APP);
    if (k < SHD) return __ldg(w + k);
    if (k < SHD + APP) return __ldg(w + SHD + GEO + (k - SHD));
    if (k < SHD + APP + GEO) return __ldg(w + SHD + (k - SHD - APP));
    return 0.f;
  });
  stage_weight<kT, 64, 64>(smem + OFF_WC1, [&](int n, int k) { return __ldg(P.col_w[1] + n * H + k); });
  stage_weight<kT, 16, 64>(smem + OFF_WC2, [&](int n, int k) { return n < 3 ? __ldg(P.col_w[2] + n * H + k) : 0.f; });
  for (int i = tid; i < F_COUNT; i += kT) {
    float v = 0.f;
    if (i < F_B1) v = __ldg(P.base_b[0] + i);
    else if (i < F_BS0) v = __ldg(P.base_b[1] + (i - F_B1));
    else if (i < F_BC0) v = __ldg(P.sem_b[0] + (i - F_BS0));
    else if (i < F_BC1) v = __ldg(P.col_b[0] + (i - F_BC0));
    else if (i < F_FOLD) v = __ldg(P.col_b[1] + (i - F_BC1));
    else if (i < F_RED) {  // fold[k] = sum_j head_w[j] * W_sem1[j][k]
      const int k = i - F_FOLD;
      float acc = 0.f;
      for (int j = 0; j < H; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_w[1] + j * H + k), acc);
      v = acc;
    }
    sf[i] = v;
  }
  if (tid < 128) {  // ONES tile: column 0 = 1 (bf16 0x3F80), everything else 0
    uint8_t* ones = smem + OFF_ONES;
    *reinterpret_cast<uint4*>(ones + row * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ones + 2048 + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = s_tmem_base;
  const uint32_t trow = tb + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* tH = smem + OFF_H;
  uint8_t* tD16 = smem + OFF_D16;
  uint8_t* tCIN = smem + OFF_CIN;
  uint8_t* tA = smem + OFF_A;
  uint8_t* tC1 = smem + OFF_C1;
  uint8_t* tDY = smem + OFF_DY;
  const uint32_t sb = smem_u32(smem);
  const uint32_t aH = sb + OFF_H, aD16 = sb + OFF_D16, aCIN = sb + OFF_CIN, aA = sb + OFF_A, aC1 = sb + OFF_C1, aDY = sb + OFF_DY,
                 aONES = sb + OFF_ONES;
  constexpr uint32_t LO64 = 128 * 64 * 2, LO32 = 128 * 32 * 2, LO16 = 128 * 16 * 2;
  const uint32_t aGEO = aCIN + GEO_CHUNK;  // [geo | 1] as a K=16 tile: hi at aGEO, lo at aGEO + LO64
  uint32_t phase = 0;

  const long long N = (long long)a.Rr.R * a.Rr.S;
  const long long tiles = (N + 127) / 128;
  const int S = a.Rr.S;
  const uint32_t hmask = (1u << F.log2T) - 1u;

  // per-thread running sums of the narrow bias gradients (half 0: do3, dlogit; half 1: dout16)
  float acc_small[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc_small[i] = 0.f;
  float acc_dlogit = 0.f;
  bool first = true;

#ifdef FNR_BWD_PROF  // role timers of CTA 0 (experiment builds only: tools/r2/build_variant.sh prof fnr_tc_bwd.cu -DFNR_BWD_PROF)
  const bool prof = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 8);
  long long t_mark = 0, t_total = clock64(), t_epi = 0, t_bar = 0, t_wait[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_empty = 0, t_full = 0, t_agg = 0, t_direct = 0;
  int wait_idx = 0, n_tiles = 0;
#define PROF_MARK() if (prof) t_mark = clock64();
#define PROF_ADD(x) if (prof) { const long long t_now = clock64(); x += t_now - t_mark; t_mark = t_now; }
#else
#define PROF_MARK()
#define PROF_ADD(x)
#endif
#define FNR_SYNC_ISSUE(...)            \
  fence_async_smem();                  \
  fence_before_sync();                 \
  PROF_ADD(t_epi)                      \
  named_bar_sync(BAR_COMPUTE, kTC);    \
  PROF_ADD(t_bar)                      \
  if (warp == 0) {                     \
    if (elect_one_sync()) {            \
      fence_after_sync();              \
      __VA_ARGS__;                     \
      mma_commit(&s_bar);              \
    }                                  \
    __syncwarp();                      \
  }
#define FNR_WAIT()          \
  PROF_ADD(t_epi)           \
  mbar_wait(&s_bar, phase); \
  phase ^= 1;               \
  fence_after_sync();       \
  PROF_ADD(t_wait[wait_idx]) \
  PROF_WAIT_NEXT()
#ifdef FNR_BWD_PROF
#define PROF_WAIT_NEXT() wait_idx = wait_idx == 8 ? 0 : wait_idx + 1;
#else
#define PROF_WAIT_NEXT()
#endif

  float* stage = reinterpret_cast<float*>(tC1);  // [128][STAGE_STRIDE]: denc[32], pos xyz, valid (see STAGE_STRIDE)

  if (!is_compute) {
    // ================= scatter warps: hash-table gradient reds, decoupled from the tensor chain =================
    reg_dec<kScatterRegs>();
    float2* const gtab = reinterpret_cast<float2*>(G.hash_table);
    const bool do_scatter = !(a.debug_flags & 1);
    named_bar_arrive(BAR_EMPTY, kT);
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      PROF_MARK()
      named_bar_sync(BAR_FULL, kT);
      PROF_ADD(t_full)
      float g[32];
      const float4* src = reinterpret_cast<const float4*>(stage + row * STAGE_STRIDE);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = src[q];
        g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
      }
      const float4 pv = src[8];
      if (tile + gridDim.x < tiles) named_bar_arrive(BAR_EMPTY, kT);  // staging may be overwritten
      const Vec3 pos = {pv.x, pv.y, pv.z};
      const bool live = pv.w != 0.f;
      // 8 scatter warps: warps 0-3 take levels 0-7 of their point, warps 4-7 levels 8-15
      const int lsplit = kScatterWarps == 8 ? ((tid - kTC) >> 7) : 0;
      // Levels 0 .. kMergeEnd-1: the 32 lanes of a warp are consecutive samples of a ray and share grid cells (run length ~25 at
      // level 0, ~2.6 at level 7 on the bench workload).  ONE shuffle round: the lane at an even position of a run of same-cell
      // lanes absorbs its successor (16 shuffles), which halves the red lanes of these levels.  Measured against the full
      // segmented scan of round 1 (5 rounds, one issuing lane per run), two rounds, and no merging in
      // profiles/r2_backward_experiments.md: this is the fastest schedule here.
      if (do_scatter && lsplit == 0) {
#pragma unroll 1
        for (int l = 0; l < kMergeEnd; ++l) scatter_level_merged1(gtab, pos, live, g[2 * l], g[2 * l + 1], l, F.scalings[l], F.log2T, hmask, lane);
      }
      PROF_ADD(t_agg)
      if (do_scatter) {
        if (live) {
#pragma unroll
          for (int l = kMergeEnd; l < 16 - kOffload; ++l) {
            const float g0 = g[2 * l], g1 = g[2 * l + 1];
            if ((kScatterWarps == 4 || (l < 8) == (lsplit == 0)) && (g0 != 0.f || g1 != 0.f)) {
              const LevelCell c = level_cell(pos, F.scalings[l]);
              const uint32_t base = (uint32_t)l << F.log2T;
              // x-neighbours (floor x even, ceil = floor + 1) are adjacent table rows r, r^1: one 16-byte red
              // instead of two 8-byte ones.  Corner pairs (x=floor, x=ceil) per (y,z): (6,5) (7,4) (2,1) (3,0).
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
          }
        }
      }
      PROF_ADD(t_direct)
    }
#ifdef FNR_BWD_PROF
    if (prof) printf("bwd scatter warp: total %lld cycles; waiting for a hand-off %lld, aggregated levels %lld, direct levels %lld\n", clock64() - t_total, t_full, t_agg, t_direct);
#endif
  } else {
  // ================= compute warps: recompute + dX / dW chain on the tensor cores =================
  reg_inc<kComputeRegs>();
  // previous tile's finest levels, scattered by warps 4-7 inside this tile's MMA waits (kOffload > 0)
  float hold[kOffload > 0 ? 2 * kOffload : 1];
  Vec3 hold_pos = {0.f, 0.f, 0.f};
  bool hold_live = false;
  float2* const gtab_c = reinterpret_cast<float2*>(G.hash_table);
#define FNR_OFFLOAD_LEVEL(i)                                                                                                        \
  if (kOffload > (i) && half == 1 && hold_live && !(a.debug_flags & 1))                                                             \
    scatter_level_direct(gtab_c, hold_pos, hold[2 * (i)], hold[2 * (i) + 1], 16 - kOffload + (i), F.scalings[16 - kOffload + (i)], F.log2T, hmask);
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    PROF_MARK()
    const long long p = tile * 128 + row;
    const bool valid = p < N;
    const long long pc = valid ? p : N - 1;
    const int ray = (int)(pc / S);
    const float* o = a.Rr.origins + 3 * (size_t)ray;
    const float* d = a.Rr.directions + 3 * (size_t)ray;
    bool sel;
    const Vec3 pos = field_position(o, d, __ldg(a.Rr.starts + pc), __ldg(a.Rr.ends + pc), F.position_mode, F.aabb, sel);
    const float vm = valid ? 1.f : 0.f;
    const float* pg = a.point_grads + 5 * (size_t)pc;
    const float d_sigma = __ldg(pg) * vm;
    const float d_logit = __ldg(pg + 4) * vm;
    const float4* stash4 = reinterpret_cast<const float4*>(a.stash_encoding + (size_t)pc * ENC);

    // ---- T0: encoding tile (from the forward's stash) -> A ; base0 ------------------------------------
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = 2 * half + jj;
      const float4 u = __ldg(stash4 + 2 * j), w = __ldg(stash4 + 2 * j + 1);
      const float c[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
      store_chunk(tA, 128 * 32 * 2, row, j, c);
    }
    FNR_SYNC_ISSUE(issue_gemm<32, 64>(tb + C_R0, aA, sb + OFF_W0))

    // ---- T1: colour-input chunks [sh | app] ; h1 -> H ; base1 ------------------------------------------
    int cam = 0;
    if (F.appearance_mode == FNR_APP_PER_CAMERA) cam = __ldg(a.Rr.camera_indices + ray);
    if (half == 0) {
      float sh[SHD];
      sh_degree4(__ldg(d), __ldg(d + 1), __ldg(d + 2), sh);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float c[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) c[q] = sh[8 * j + q];
        store_chunk(tCIN, 128 * 64 * 2, row, j, c);
      }
    } else {
      const float4* e4 = reinterpret_cast<const float4*>(P.app_embedding + (size_t)cam * APP);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (F.appearance_mode == FNR_APP_PER_CAMERA) {
          const float4 u = __ldg(e4 + 2 * j), w = __ldg(e4 + 2 * j + 1);
          c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w; c[4] = w.x; c[5] = w.y; c[6] = w.z; c[7] = w.w;
        }
        store_chunk(tCIN, 128 * 64 * 2, row, 2 + j, c);
      }
    }
    FNR_WAIT()
    epi32(trow + C_R0 + 32 * half, tH, row, half, [&](int n, float x) { return fmaxf(x + sf[F_B0 + n], 0.f); });
    FNR_SYNC_ISSUE(issue_gemm<64, 16>(tb + C_R1, aH, sb + OFF_W1))

    FNR_OFFLOAD_LEVEL(6)
    // ---- T2: [h0 | geo] ; geo chunks ; dlogit tile ; semantic0 + colour0 --------------------------------
    FNR_WAIT()
    float d_h0;
    {
      uint32_t r0[16];
      tmem_ld16(trow + C_R1, r0);
      tmem_ld_wait();
      float outv[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) outv[n] = __uint_as_float(r0[n]) + sf[F_B1 + n];
      // density = exp(h0) * selector ; trunc_exp backward: g * exp(clamp(h0, -15, 15))
      d_h0 = sel ? d_sigma * expf(fminf(fmaxf(outv[0], -15.f), 15.f)) : 0.f;
      float g[8];
      if (half == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = outv[1 + q];
        store_chunk(tCIN, 128 * 64 * 2, row, 6, g);
        const float dl0[8] = {d_logit, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(tD16, 128 * 16 * 2, row, 0, dl0);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = q < 7 ? outv[9 + q] : 1.0f;  // column 63 (= 15 of [geo|1]) = 1: bias-gradient column
        store_chunk(tCIN, 128 * 64 * 2, row, 7, g);
        const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(tD16, 128 * 16 * 2, row, 1, z8);
      }
    }
    FNR_SYNC_ISSUE(issue_gemm_lo<16, 64>(tb + C_R2, aGEO, aGEO + LO64, sb + OFF_WS0); issue_gemm<64, 64>(tb + C_R0, aCIN, sb + OFF_WC0))

    FNR_OFFLOAD_LEVEL(7)
    // ---- T3: z1 -> A, dz1 -> DY, c1 -> C1 ; colour1 + AV + AS0 ----------------------------------------
    FNR_WAIT()
    {
      uint32_t r[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld8(trow + C_R2 + 32 * half + 8 * j, r[j]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z[8], dz[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = 32 * half + 8 * j + q;
          z[q] = fmaxf(__uint_as_float(r[j][q]) + sf[F_BS0 + n], 0.f);
          dz[q] = z[q] > 0.f ? d_logit * sf[F_FOLD + n] : 0.f;
        }
        store_chunk(tA, 128 * 64 * 2, row, 4 * half + j, z);
        store_chunk(tDY, 128 * 64 * 2, row, 4 * half + j, dz);
      }
    }
    PROF_ADD(t_epi)
    named_bar_sync(BAR_EMPTY, kT);  // the scatter warps have copied the previous tile's hand-off out of the C1 region
    PROF_ADD(t_empty)
    epi32(trow + C_R0 + 32 * half, tC1, row, half, [&](int n, float x) { return fmaxf(x + sf[F_BC0 + n], 0.f); });
    FNR_SYNC_ISSUE(issue_gemm<64, 64>(tb + C_R0, aC1, sb + OFF_WC1); issue_dw<16>(tb + C_AV, aA, aA + LO64, aD16, aD16 + LO16, !first);
                   issue_dw<16>(tb + C_AS0, aDY, aDY + LO64, aGEO, aGEO + LO64, !first))

    FNR_OFFLOAD_LEVEL(0)
    // ---- T4: c2 -> A ; do3 -> D16 ; AC2 + dc2 -----------------------------------------------------------
    FNR_WAIT()
    epi32(trow + C_R0 + 32 * half, tA, row, half, [&](int n, float x) { return fmaxf(x + sf[F_BC1 + n], 0.f); });
    if (half == 0) {
      const float* rgb = a.sample_rgb + 3 * (size_t)pc;
      float do3[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float c = __ldg(rgb + i);
        do3[i] = __ldg(pg + 1 + i) * vm * c * (1.0f - c);
        acc_small[i] += do3[i];
      }
      acc_dlogit += d_logit;
      const float c0[8] = {do3[0], do3[1], do3[2], 0.f, 0.f, 0.f, 0.f, 0.f};
      store_chunk(tD16, 128 * 16 * 2, row, 0, c0);
    } else {
      const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store_chunk(tD16, 128 * 16 * 2, row, 1, z8);
    }
    FNR_SYNC_ISSUE(issue_dw<16>(tb + C_AC2, aA, aA + LO64, aD16, aD16 + LO16, !first); issue_dx<16, 64>(tb + C_R2, aD16, sb + OFF_WC2))

    FNR_OFFLOAD_LEVEL(4)
    // ---- T5: dc2 = (do3 Wc2) * relu'(c2) -> DY ; AC1, AC1b, dc1 ----------------------------------------
    FNR_WAIT()
    epi32_masked(trow + C_R2 + 32 * half, tA, tDY, row, half);
    FNR_SYNC_ISSUE(issue_dw<64>(tb + C_AC1, aDY, aDY + LO64, aC1, aC1 + LO64, !first); issue_dw<16>(tb + C_AC1B, aDY, aDY + LO64, aONES, 0u, !first);
                   issue_dx<64, 64>(tb + C_R0, aDY, sb + OFF_WC1))

    FNR_OFFLOAD_LEVEL(1)
    // ---- T6: dc1 = (dc2 Wc1) * relu'(c1) -> DY ; AC0, dcin --------------------------------------------
    FNR_WAIT()
    epi32_masked(trow + C_R0 + 32 * half, tC1, tDY, row, half);
    FNR_SYNC_ISSUE(issue_dw<64>(tb + C_AC0, aDY, aDY + LO64, aCIN, aCIN + LO64, !first); issue_dx<64, 64>(tb + C_R2, aDY, sb + OFF_WC0))

    FNR_OFFLOAD_LEVEL(2)
    // ---- T7: dcin -> d_app (embedding gradient), d_geo ; dout16 -> D16 ; AB1, dh1 -----------------------
    FNR_WAIT()
    {
      // half 0: columns 16..31 = d_app[0..15];  half 1: columns 32..47 = d_app[16..31], 48..62 = d_geo
      uint32_t ra[16], rg[16];
      tmem_ld16(trow + C_R2 + 16 + 16 * half, ra);
      if (half == 1) tmem_ld16(trow + C_R2 + 48, rg);
      tmem_ld_wait();
      if (F.appearance_mode == FNR_APP_PER_CAMERA) {
        const bool uniform = __all_sync(kTcFullMask, cam == __shfl_sync(kTcFullMask, cam, 0));
        if (uniform) {
          // transposed butterfly over 16 values: after offsets 16,8,4,2 lane L holds (half-)sums of value L>>1
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = __uint_as_float(ra[i]);
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
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = __uint_as_float(ra[i]);
            if (v != 0.f) atomicAdd(G.app_embedding + (size_t)cam * APP + 16 * half + i, v);
          }
        }
      }
      if (half == 1) {
        float c0[8], c1[8];
        c0[0] = d_h0;
#pragma unroll
        for (int q = 1; q < 8; ++q) c0[q] = __uint_as_float(rg[q - 1]);
#pragma unroll
        for (int q = 0; q < 8; ++q) c1[q] = __uint_as_float(rg[7 + q]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc_small[q] += c0[q];
          acc_small[8 + q] += c1[q];
        }
        store_chunk(tD16, 128 * 16 * 2, row, 0, c0);
        store_chunk(tD16, 128 * 16 * 2, row, 1, c1);
      }
    }
    FNR_SYNC_ISSUE(issue_dw<16>(tb + C_AB1, aH, aH + LO64, aD16, aD16 + LO16, !first); issue_dx<16, 64>(tb + C_R0, aD16, sb + OFF_W1))

    FNR_OFFLOAD_LEVEL(5)
    // ---- T8: dh1 = (dout W1) * relu'(h1) -> DY ; reload enc -> A ; AB0, AB0b, denc ------------------------
    FNR_WAIT()
    epi32_masked(trow + C_R0 + 32 * half, tH, tDY, row, half);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = 2 * half + jj;
      const float4 u = __ldg(stash4 + 2 * j), w = __ldg(stash4 + 2 * j + 1);
      const float c[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
      store_chunk(tA, 128 * 32 * 2, row, j, c);
    }
    FNR_SYNC_ISSUE(issue_dw<32>(tb + C_AB0, aDY, aDY + LO64, aA, aA + LO32, !first); issue_dw<16>(tb + C_AB0B, aDY, aDY + LO64, aONES, 0u, !first);
                   issue_dx<64, 32>(tb + C_R2, aDY, sb + OFF_W0))

    FNR_OFFLOAD_LEVEL(3)
    // ---- T9: denc -> hand-off to the scatter warps (C1 region is idle since T6) ------------------------------
    FNR_WAIT()
    {
      uint32_t r0[16];
      tmem_ld16(trow + C_R2 + 16 * half, r0);
      tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_float4(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1]), __uint_as_float(r0[4 * q + 2]), __uint_as_float(r0[4 * q + 3]));
      if (half == 0) reinterpret_cast<float4*>(stage + row * STAGE_STRIDE + 32)[0] = make_float4(pos.x, pos.y, pos.z, valid ? 1.f : 0.f);
      if (kOffload > 0 && half == 1) {  // columns 32 - 2 kOffload .. 31 = levels 16 - kOffload .. 15
#pragma unroll
        for (int i = 0; i < 2 * kOffload; ++i) hold[i] = __uint_as_float(r0[16 - 2 * kOffload + i]);
        hold_pos = pos;
        hold_live = valid;
      }
    }
    named_bar_arrive(BAR_FULL, kT);
    PROF_ADD(t_epi)
#ifdef FNR_BWD_PROF
    ++n_tiles;
#endif
    fence_before_sync();  // this tile's TMEM reads are ordered before the next tile's MMAs (via the next barrier)
    first = false;
  }
  FNR_OFFLOAD_LEVEL(0) FNR_OFFLOAD_LEVEL(1) FNR_OFFLOAD_LEVEL(2) FNR_OFFLOAD_LEVEL(3) FNR_OFFLOAD_LEVEL(4) FNR_OFFLOAD_LEVEL(5) FNR_OFFLOAD_LEVEL(6) FNR_OFFLOAD_LEVEL(7)  // the last tile's
#undef FNR_OFFLOAD_LEVEL
#undef FNR_SYNC_ISSUE
#undef FNR_WAIT
#ifdef FNR_BWD_PROF
  if (prof)
    printf("bwd compute warp %d: %d tiles, total %lld cycles; epilogues + loads %lld, compute barrier %lld, waiting for the scatter warps %lld, "
           "MMA waits T1..T9 %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", warp, n_tiles, clock64() - t_total, t_epi, t_bar, t_empty, t_wait[0], t_wait[1],
           t_wait[2], t_wait[3], t_wait[4], t_wait[5], t_wait[6], t_wait[7], t_wait[8]);
#endif
  }  // compute warps

  // ---- flush the resident accumulators ---------------------------------------------------------------
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const bool did_work = (long long)blockIdx.x < tiles;
  if (did_work && is_compute) {
    // M=64 accumulators: row i lives in TMEM lane (i % 16) + 32 * (i / 16)  ->  owner threads: lane < 16,
    // i = 16 * (warp & 3) + lane; the two halves split the columns of every accumulator.
    const bool owner = lane < 16;
    const int i = 16 * (warp & 3) + lane;
    {  // AV: v[i] = sum_points z1[i] * dlogit -> shared scratch ;  AS0 ; AC2 ; AB1   (16-column accumulators)
      uint32_t q[8];
      tmem_ld8(trow + C_AV, q);
      tmem_ld_wait();
      if (owner && half == 0) sf[F_RED + i] = __uint_as_float(q[0]);
      tmem_ld8(trow + C_AS0 + 8 * half, q);  // d W_sem0[n=i][k<15], column 15 = d b_sem0[i]
      tmem_ld_wait();
      if (owner) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = 8 * half + kk;
          if (k < GEO) atomicAdd(G.sem_w[0] + i * GEO + k, __uint_as_float(q[kk]));
          else atomicAdd(G.sem_b[0] + i, __uint_as_float(q[kk]));
        }
      }
      tmem_ld8(trow + C_AC2, q);  // rows = c2 feature k=i, cols = colour output n<3
      tmem_ld_wait();
      if (owner && half == 0) {
#pragma unroll
        for (int n = 0; n < 3; ++n) atomicAdd(G.col_w[2] + n * H + i, __uint_as_float(q[n]));
      }
      tmem_ld8(trow + C_AB1 + 8 * half, q);  // rows = h1 feature k=i, cols = base output n<16
      tmem_ld_wait();
      if (owner) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) atomicAdd(G.base_w[1] + (8 * half + kk) * H + i, __uint_as_float(q[kk]));
      }
      tmem_ld8(trow + C_AC1B, q);
      tmem_ld_wait();
      if (owner && half == 0) atomicAdd(G.col_b[1] + i, __uint_as_float(q[0]));
      tmem_ld8(trow + C_AB0B, q);
      tmem_ld_wait();
      if (owner && half == 1) atomicAdd(G.base_b[0] + i, __uint_as_float(q[0]));
    }
    uint32_t r[32];
    // AC1: d W_col1[n=i][k]
    tmem_ld32(trow + C_AC1 + 32 * half, r);
    tmem_ld_wait();
    if (owner) {
#pragma unroll
      for (int k = 0; k < 32; ++k) atomicAdd(G.col_w[1] + i * H + 32 * half + k, __uint_as_float(r[k]));
    }
    // AC0: cols in [sh | app | geo | bias] order -> torch order [sh | geo | app]
    tmem_ld32(trow + C_AC0 + 32 * half, r);
    tmem_ld_wait();
    if (owner) {
      float* gw = G.col_w[0] + i * (SHD + GEO + APP);
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const int k = 32 * half + kk;
        const float val = __uint_as_float(r[kk]);
        if (k < SHD) atomicAdd(gw + k, val);
        else if (k < SHD + APP) atomicAdd(gw + SHD + GEO + (k - SHD), val);
        else if (k < SHD + APP + GEO) atomicAdd(gw + SHD + (k - SHD - APP), val);
        else atomicAdd(G.col_b[0] + i, val);
      }
    }
    // AB0: d W_base0[n=i][k<32]
    {
      uint32_t q[16];
      tmem_ld16(trow + C_AB0 + 16 * half, q);
      tmem_ld_wait();
      if (owner) {
#pragma unroll
        for (int k = 0; k < 16; ++k) atomicAdd(G.base_w[0] + i * ENC + 16 * half + k, __uint_as_float(q[k]));
      }
    }
    // narrow bias gradients + sum of dlogit: warp reduce, then one atomic per warp
    if (half == 0) {
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const float s = warp_sum_f(acc_small[n]);
        if (lane == 0) atomicAdd(G.col_b[2] + n, s);
      }
      const float sdl_w = warp_sum_f(acc_dlogit);
      if (lane == 0) sf[F_RED + 64 + warp] = sdl_w;
    } else {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const float s = warp_sum_f(acc_small[n]);
        if (lane == 0) atomicAdd(G.base_b[1] + n, s);
      }
    }
  }
  __syncthreads();
  if (did_work) {
    // folded semantic tail: logit = head_w . (W1 z1 + b1) + head_b, with v = sum dlogit * z1, s = sum dlogit
    const float* v = sf + F_RED;
    float sdl = 0.f;
    for (int w = 0; w < 4; ++w) sdl += sf[F_RED + 64 + w];
    if (tid == 0) atomicAdd(G.head_b, sdl);
    if (tid < H) {
      const int j = tid;
      const float hw = __ldg(P.head_w + j);
      atomicAdd(G.sem_b[1] + j, hw * sdl);
      float acc = __ldg(P.sem_b[1] + j) * sdl;
      for (int k = 0; k < H; ++k) acc = fmaf(__ldg(P.sem_w[1] + j * H + k), v[k], acc);
      atomicAdd(G.head_w + j, acc);
    }
    for (int e = tid; e < H * H; e += kT) {
      const int j = e / H, k = e % H;
      atomicAdd(G.sem_w[1] + e, __ldg(P.head_w + j) * v[k]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

}  // namespace

bool tc_backward_supported(Family fam, const KField& F, const KRays& Rr, const KFieldBwd& B) {
  (void)Rr;
  return fam == kFamilySmall && B.stash_encoding != nullptr && B.sample_rgb != nullptr && !F.pass_semantic_gradients &&
         (F.appearance_mode == FNR_APP_PER_CAMERA || F.appearance_mode == FNR_APP_ZEROS);
}

int launch_tc_field_backward(Family fam, const KField& F, const KParams& P, const KParams& G, const KRays& Rr, const KFieldBwd& B,
                             cudaStream_t st) {
  if (!tc_backward_supported(fam, F, Rr, B)) {
    set_error("tcgen05 backward kernel does not support this configuration");
    return FNR_ERR_UNSUPPORTED;
  }
  const long long N = (long long)Rr.R * Rr.S;
  if (N == 0) return FNR_OK;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_field_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_field_backward_kernel)");
    configured = true;
  }
  BwdArgs a;
  a.F = F;
  a.P = P;
  a.G = G;
  a.Rr = Rr;
  a.point_grads = B.point_grads;
  a.stash_encoding = B.stash_encoding;
  a.sample_rgb = B.sample_rgb;
  {
    static int flags = -1;
    if (flags < 0) {
      const char* e = getenv("FNR_DEBUG_BWD");
      flags = e ? atoi(e) : 0;
      const int skip = (flags >> 1) & 1;
      cudaMemcpyToSymbol(g_skip_dw, &skip, sizeof(int));
    }
    a.debug_flags = flags;
  }
  const long long tiles = (N + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc_field_backward_kernel<<<grid, kT, kSmem, st>>>(a);
  return check_launch("tc_field_backward_kernel");
}

}  // namespace fnr

// Warp-specialised fused render / export forward of the fruit_nerf family for sm_100a
// (fruit_nerf/fruit_field.py:168-301 + fruit_nerf/fruit_nerf.py:320-348 / 251-269 in one pass).
//
// One persistent CTA per SM, 24 warps in three roles that only meet through mbarriers:
//   * 8 GATHER warps (thread pair (row, hf) per point: hash levels 8 hf .. 8 hf + 7).  They never touch the MLPs: per
//     (tile, level) item a thread computes the cell, issues the corner rows as `cp.async` (LDGSTS: 4 x 16 B x-neighbour
//     pairs + up to 4 x 8 B unpaired ceil-x rows) into a private shared-memory stage and, kDepth items later,
//     blends the item out of shared memory.  Rows in flight cost shared memory, not registers, and the stream of
//     gathers never stops for a tensor-core phase: the L1TEX row rate is the only thing that paces these warps.
//     A finished tile's 32 features go straight into TENSOR MEMORY (tcgen05.st, bf16 hi/lo split) as the A operand
//     of the first GEMM -- a 3-deep ring of encoding buffers decouples the gather from the MLP chains -- and into
//     the encoding stash of the backward (one 32-byte sector per store).
//   * 2 CHAIN slots of 8 warps (thread pair (row, half): accumulator columns 32 half ..).  Every activation lives in
//     tensor memory: an epilogue reads 16 accumulator columns (tcgen05.ld), applies bias / ReLU, re-splits to
//     bf16 hi/lo and writes the next layer's A operand IN PLACE over the columns it just read (tcgen05.st); the MMAs
//     take A from TMEM (tcgen05.mma [d], [a], b-desc) and the weights from shared memory.  Five round trips per tile:
//     base0 -> base1 -> [semantic0 || colour0] -> [folded semantic head || colour1] -> colour2.
//   * the chain warps of both slots composite (or threshold + compact, export) a group of whole rays from shared memory.
// Operand layout in TMEM ("k-step interleaved"): K elements [16 s, 16 s + 16) of a row sit in columns
// [16 s, 16 s + 8) (hi halves, two bf16 per column) and [16 s + 8, 16 s + 16) (lo halves): a 16-column accumulator
// chunk is replaced by the 16 columns of the operand chunk made from it.
// TMEM columns: slot s at 208 s: X [0,64) Y [64,80) C [80,144) Z [144,208); encoding ring at 416 + 32 e, e < 3.
#include <cstdio>
#include <cstdlib>
#include "fnr_common.cuh"
#include "fnr_kernels.h"
#include "fnr_tcgen05.cuh"
#include "fnr_tc_common.cuh"
#include "fnr_tc_group.cuh"

namespace fnr {
using namespace tc;
using namespace tcx;

namespace {

constexpr int kSlots = 2;
constexpr int kSlotThreads = 256;
constexpr int kChainThreads = kSlots * kSlotThreads;  // warps 0..15
#ifndef FNR_WS_GATHER_WARPS
#define FNR_WS_GATHER_WARPS 16
#endif
constexpr int kGatherWarps = FNR_WS_GATHER_WARPS;     // 8: thread pair per point (8 levels each); 16: four threads per point (4 levels each)
constexpr int kGatherThreads = kGatherWarps * 32;     // warps 16..
constexpr int kLv = 16 / (kGatherWarps / 4);          // hash levels per gather thread
constexpr int kCoarseLevels = 8;                      // levels 0..7 are loaded directly (coalescing), 8..15 staged with cp.async
constexpr int kCtaThreads = kChainThreads + kGatherThreads;
constexpr int kMaxGroupPoints = 768;
constexpr int kEnc = 3;    // encoding buffers in tensor memory
#ifndef FNR_WS_DEPTH
#define FNR_WS_DEPTH 2
#endif
constexpr int kDepth = FNR_WS_DEPTH;  // (tile, level) items in flight per gather thread

constexpr int GEO = 15, ENC = 32, H = 64, APP = 32, SHD = 16;
constexpr int K_BASE0 = 32, N_BASE0 = 64;
constexpr int K_BASE1 = 64, N_BASE1 = 16;   // outputs [h0 | geo 0..14]
constexpr int K_SEM0 = 16, N_SEM0 = 64;     // K order [h0 slot (zero weights) | geo 0..14]
constexpr int K_SEMH = 64, N_SEMH = 16;     // folded (semantic layer 1) x head -> row 0
constexpr int K_COL0 = 64, N_COL0 = 64;     // K order [sh 16 | app 32 | h0 slot | geo 0..14]
constexpr int K_COL1 = 64, N_COL1 = 64;
constexpr int K_COL2 = 64, N_COL2 = 16;

constexpr int OFF_W_BASE0 = 0;
constexpr int OFF_W_BASE1 = OFF_W_BASE0 + 2 * wbytes(N_BASE0, K_BASE0);
constexpr int OFF_W_SEM0 = OFF_W_BASE1 + 2 * wbytes(N_BASE1, K_BASE1);
constexpr int OFF_W_SEMH = OFF_W_SEM0 + 2 * wbytes(N_SEM0, K_SEM0);
constexpr int OFF_W_COL0 = OFF_W_SEMH + 2 * wbytes(N_SEMH, K_SEMH);
constexpr int OFF_W_COL1 = OFF_W_COL0 + 2 * wbytes(N_COL0, K_COL0);
constexpr int OFF_W_COL2 = OFF_W_COL1 + 2 * wbytes(N_COL1, K_COL1);
constexpr int OFF_BIAS = OFF_W_COL2 + 2 * wbytes(N_COL2, K_COL2);
constexpr int B_BASE0 = 0, B_BASE1 = 64, B_SEM0 = 80, B_SEMH = 144, B_COL0 = 160, B_COL1 = 224, B_COL2 = 288, B_APP = 304,
              B_COUNT = 336;
constexpr int OFF_SAMPLES = OFF_BIAS + B_COUNT * 4;                 // [kMaxGroupPoints][5] floats
constexpr int OFF_SEL = OFF_SAMPLES + kMaxGroupPoints * 5 * 4;      // [kEnc][128] selector bytes
// gather stage of one warp and one item: 4 x [lane] 16 B pair rows | 4 x [lane] 8 B single rows | [lane] 16 B (ox, oy, oz, flags)
constexpr int STAGE_PAIR = 0, STAGE_SINGLE = 4 * 512, STAGE_META = STAGE_SINGLE + 4 * 256, STAGE_BYTES = STAGE_META + 512;
constexpr int OFF_STAGE = (OFF_SEL + kEnc * 128 + 127) & ~127;
constexpr int kFineWarps = kGatherWarps / 2;  // the gather warps of levels 8..15 stage their rows with cp.async; the others load directly
constexpr int OFF_END = OFF_STAGE + kFineWarps * kDepth * STAGE_BYTES;
constexpr int kSmemBytes = OFF_END + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared-memory budget");
static_assert(OFF_BIAS % 16 == 0 && OFF_SAMPLES % 16 == 0 && STAGE_BYTES % 128 == 0, "alignment");
static_assert(kDepth >= 1 && kDepth <= kLv, "pipeline depth");
static_assert(kGatherWarps == 8 || kGatherWarps == 16, "gather warps");

constexpr int C_X = 0, C_Y = 64, C_C = 80, C_Z = 144, kSlotCols = 208, C_ENC = kSlots * kSlotCols;
static_assert(C_ENC + kEnc * 32 <= 512, "tensor-memory budget");

constexpr int BAR_SLOT0 = 1, BAR_CHAIN = 3;
// register split (setmaxnreg): 768 threads are launched with 80 registers; the gather warps hand 16 each to the chain warps
constexpr int kLaunchRegs = kCtaThreads == 1024 ? 64 : 80;
constexpr int kGatherRegs = kCtaThreads == 1024 ? 56 : 64, kChainRegs = kCtaThreads == 1024 ? 72 : 88;
static_assert(kGatherThreads * kGatherRegs + kChainThreads * kChainRegs <= kCtaThreads * kLaunchRegs, "register pool");

struct WsArgs {
  KField F;
  KParams P;
  KRays Rr;
  KFieldOut O;
  KComposite Cm;
  int rays_per_group;
  int composite;
  int debug;  // FNR_DEBUG_FWD (timing experiments only; results are wrong with bits 2..8 set): bit 1 block 0 prints where each role
              // waited, 2 fine warps skip their table copies, 4 coarse warps skip their table loads, 8 chain slots skip the MLPs
  KExport E;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

// K elements [16 s, 16 s + 16) of this thread's row -> operand region `reg` (lane-adjusted address): [hi 8 | lo 8] columns
__device__ __forceinline__ void st_kstep(uint32_t reg, int s, const float (&v)[16]) {
  uint32_t w[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], w[q], w[8 + q]);
  tmem_st16(reg + 16 * s, w);
}

// a gather thread's share of the encoding operand (K = 32, two k-steps): 8 levels = k-step lq; 4 levels = half of k-step lq / 2
__device__ __forceinline__ void st_encoding(uint32_t reg, int lq, const float (&v)[16]) { st_kstep(reg, lq, v); }
__device__ __forceinline__ void st_encoding(uint32_t reg, int lq, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], h[q], l[q]);
  const uint32_t col = reg + 16 * (lq >> 1) + 4 * (lq & 1);  // K elements 8 lq .. 8 lq + 7 -> 4 hi + 4 lo columns
  tmem_st4(col, h);
  tmem_st4(col + 8, l);
}

// in-place epilogue of accumulator columns [16 s, 16 s + 16): v = f(column, x) replaces them as operand k-step s
template <class Fn>
__device__ __forceinline__ void epi_inplace16(uint32_t reg, int s, Fn f) {
  uint32_t r[16];
  tmem_ld16(reg + 16 * s, r);
  tmem_ld_wait();
  float v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = f(16 * s + q, __uint_as_float(r[q]));
  st_kstep(reg, s, v);
}

// D[128,N] = A[128,K] (tensor memory, k-step interleaved hi/lo) x W[N,K]^T (shared memory, hi/lo): 3 MMAs per K step
template <int K, int N>
__device__ __forceinline__ void issue_gemm_tsi(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_hi) {
  constexpr uint32_t idesc = idesc_bf16_f32(128, N);
  constexpr uint32_t w_lo_off = N * K * 2;
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t wh = smem_desc(w_hi + ks * 2 * N * 16, N * 16, 128);
    const uint64_t wl = smem_desc(w_hi + w_lo_off + ks * 2 * N * 16, N * 16, 128);
    mma_ts(d_tmem, a_tmem + 16 * ks, wh, idesc, ks > 0);
    mma_ts(d_tmem, a_tmem + 16 * ks + 8, wh, idesc, true);
    mma_ts(d_tmem, a_tmem + 16 * ks, wl, idesc, true);
  }
}

// waits that are expected to be long (another role has to produce a tile): back off so that the spin does not take issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
#ifdef FNR_WS_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// predicated form (no divergent branch around the copy)
__device__ __forceinline__ void cp_async8_if(bool pred, uint32_t dst, const void* src) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}" ::"r"(dst), "l"(src), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// the point a thread serves in a tile: ray, interval -> position (bit-identical to the oracle), selector
struct TilePoint {
  bool valid, sel;
  int local, ray;
  size_t gp;
  Vec3 pos;
};

template <bool kExport>
__device__ __forceinline__ TilePoint tile_point(const WsArgs& a, int ray0, int pts, int S, int tile, int row, bool need_pos) {
  TilePoint t;
  t.local = tile * 128 + row;
  t.valid = t.local < pts;
  const int lc = t.valid ? t.local : pts - 1;
  t.ray = ray0 + lc / S;
  t.gp = (size_t)ray0 * S + lc;
  t.sel = false;
  t.pos = Vec3{0.f, 0.f, 0.f};
  if (need_pos) {
    const float* o = (kExport ? a.E.origins : a.Rr.origins) + 3 * (size_t)t.ray;
    const float* d = kExport ? a.E.normal : a.Rr.directions + 3 * (size_t)t.ray;
    float t0, t1;
    if constexpr (kExport) {
      export_interval(a.E, t.ray, lc % S, t0, t1);
    } else {
      t0 = __ldg(a.Rr.starts + t.gp);
      t1 = __ldg(a.Rr.ends + t.gp);
    }
    t.pos = field_position(o, d, t0, t1, a.F.position_mode, a.F.aabb, t.sel);
  }
  return t;
}

// kDbg: the FNR_DEBUG_FWD instantiation (role timers, experiment switches); the production instantiation carries none of it
template <bool kExport, bool kDbg>
__global__ void __launch_bounds__(kCtaThreads, 1) tc_render_forward_ws_kernel(const __grid_constant__ WsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_bar[kSlots], s_full[kEnc], s_empty[kEnc];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const KParams& P = a.P;
  const KField& F = a.F;
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* s_samples = reinterpret_cast<float*>(smem + OFF_SAMPLES);
  uint8_t* s_sel = smem + OFF_SEL;

  // ---- one-time setup: TMEM, barriers, weights ------------------------------------------------
  if (warp == 0) tmem_alloc(&s_tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&s_bar[i], 1);
    for (int i = 0; i < kEnc; ++i) {
      mbar_init(&s_full[i], kGatherThreads);
      mbar_init(&s_empty[i], 1);
    }
    mbar_fence_init();
  }
  stage_weight<kCtaThreads, N_BASE0, K_BASE0>(smem + OFF_W_BASE0, [&](int n, int k) { return __ldg(P.base_w[0] + n * ENC + k); });
  stage_weight<kCtaThreads, N_BASE1, K_BASE1>(smem + OFF_W_BASE1, [&](int n, int k) { return __ldg(P.base_w[1] + n * H + k); });
  stage_weight<kCtaThreads, N_SEM0, K_SEM0>(smem + OFF_W_SEM0, [&](int n, int k) { return k >= 1 ? __ldg(P.sem_w[0] + n * GEO + (k - 1)) : 0.f; });
  // fold: logit = head_w . (W1 z + b1) + head_b  =>  row 0 of the N=16 tile is head_w^T W1
  stage_weight<kCtaThreads, N_SEMH, K_SEMH>(smem + OFF_W_SEMH, [&](int n, int k) {
    if (n != 0) return 0.f;
    float acc = 0.f;
    for (int j = 0; j < H; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_w[1] + j * H + k), acc);
    return acc;
  });
  // colour layer 0 with the K order [sh | app | h0 slot | geo] (torch order is [sh | geo | app])
  stage_weight<kCtaThreads, N_COL0, K_COL0>(smem + OFF_W_COL0, [&](int n, int k) {
    const float* w = P.col_w[0] + n * (SHD + GEO + APP);
    if (k < SHD) return __ldg(w + k);
    if (k < SHD + APP) return __ldg(w + SHD + GEO + (k - SHD));
    if (k > SHD + APP) return __ldg(w + SHD + (k - SHD - APP - 1));
    return 0.f;
  });
  stage_weight<kCtaThreads, N_COL1, K_COL1>(smem + OFF_W_COL1, [&](int n, int k) { return __ldg(P.col_w[1] + n * H + k); });
  stage_weight<kCtaThreads, N_COL2, K_COL2>(smem + OFF_W_COL2, [&](int n, int k) { return n < 3 ? __ldg(P.col_w[2] + n * H + k) : 0.f; });
  for (int i = tid; i < B_COUNT; i += kCtaThreads) {
    float v = 0.f;
    if (i < B_BASE1) v = __ldg(P.base_b[0] + i);
    else if (i < B_SEM0) v = __ldg(P.base_b[1] + (i - B_BASE1));
    else if (i < B_SEMH) v = __ldg(P.sem_b[0] + (i - B_SEM0));
    else if (i == B_SEMH) {
      float acc = __ldg(P.head_b);
      for (int j = 0; j < H; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_b[1] + j), acc);
      v = acc;
    } else if (i < B_COL0) v = 0.f;
    else if (i < B_COL1) v = __ldg(P.col_b[0] + (i - B_COL0));
    else if (i < B_COL2) v = __ldg(P.col_b[1] + (i - B_COL1));
    else if (i < B_COL2 + 3) v = __ldg(P.col_b[2] + (i - B_COL2));
    else if (i >= B_APP && F.appearance_mode == FNR_APP_MEAN) {
      float acc = 0.f;
      for (int r = 0; r < F.num_images; ++r) acc += __ldg(P.app_embedding + (size_t)r * APP + (i - B_APP));
      v = acc / (float)F.num_images;
    }
    s_bias[i] = v;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const int S = kExport ? a.E.S : a.Rr.S, R = kExport ? a.E.B : a.Rr.R;
  const int G = a.rays_per_group;
  const int num_groups = (R + G - 1) / G;
  const uint32_t tmem0 = s_tmem_base;
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;  // TMEM lane = point of the tile (warp w may touch lanes 32 (w % 4) ..)
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  const bool prof = kDbg && (a.debug & 1) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 16 || warp == 16 + kGatherWarps / 2);
  long long t_wait_a = 0, t_wait_b = 0, t_wait_c = 0, t_wait_d = 0, t_total = 0;
  if (prof) t_total = clock64();

  if (warp >= kChainThreads / 32) {
    // ========================= gather warps =========================
    reg_dec<kGatherRegs>();
    const int gw = warp - kChainThreads / 32;
    const int lq = gw >> 2;          // this warp's levels: kLv lq .. kLv lq + kLv - 1
    const int l0 = kLv * lq;
    const bool coarse = l0 < kCoarseLevels;
    const float2* __restrict__ table = reinterpret_cast<const float2*>(P.hash_table);
    const uint32_t hmask = (1u << F.log2T) - 1u;
    const uint32_t stage0 = smem_u32(smem + OFF_STAGE) + (uint32_t)((gw - 4 * (kCoarseLevels / kLv)) * kDepth) * STAGE_BYTES;  // fine warps only
    const uint32_t enc0 = tmem0 + lane_off + C_ENC;

    float enc[2 * kLv];
    int seq = 0;          // tile sequence number of the tile being gathered

    // a finished tile: stash (backward), this thread's K elements [2 l0, 2 l0 + 2 kLv) of the encoding operand into the ring
    // buffer, selector bytes, signal the chain slot
    auto finish_tile = [&](const TilePoint& tp, int tseq) {
      if (!kExport && a.O.stash_encoding && tp.valid) {
        float* stp = a.O.stash_encoding + tp.gp * ENC + 2 * l0;  // 32-byte aligned: whole sectors per thread
#pragma unroll
        for (int j = 0; j < kLv / 4; ++j)
          st_global_v8(stp + 8 * j, enc[8 * j], enc[8 * j + 1], enc[8 * j + 2], enc[8 * j + 3], enc[8 * j + 4], enc[8 * j + 5], enc[8 * j + 6], enc[8 * j + 7]);
      }
      const int e = tseq % kEnc, use = tseq / kEnc;
      long long t0 = 0;
      if (prof) t0 = clock64();
      mbar_wait_sleep(&s_empty[e], (uint32_t)((use & 1) ^ 1));  // the base0 GEMM of this buffer's previous tile has completed
      if (prof) t_wait_a += clock64() - t0;
      fence_after_sync();
      st_encoding(enc0 + 32 * e, lq, enc);
      if (lq == 0) s_sel[e * 128 + row] = tp.sel ? 1 : 0;
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(&s_full[e]);
    };

    if (coarse) {
      // ---- coarse levels: the 32 lanes of a warp are consecutive samples of a ray and share grid cells (a warp spans 1.3 cells at
      // level 0, 12 at level 7 on the bench batch), so a direct load coalesces to a few sectors per instruction while a cp.async
      // costs one L1TEX pass per LANE.  Loads go through registers.
      for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        const int ray0 = group * G;
        const int pts = min(G, R - ray0) * S;
        const int tiles = (pts + 127) / 128;
#pragma unroll 1
        for (int tile = 0; tile < tiles; ++tile, ++seq) {
          const TilePoint cur = tile_point<kExport>(a, ray0, pts, S, tile, row, true);
#pragma unroll
          for (int li = 0; li < kLv; ++li) {
            const int l = l0 + li;
            const LevelCell c = level_cell(cur.pos, F.scalings[l]);
            const uint32_t base = (uint32_t)l << F.log2T;
            // x-neighbours (floor x even, ceil = floor + 1) are adjacent table rows r, r^1: one 16-byte load fetches both; only
            // lanes with an odd floor x load the ceil-x row separately.  Corner pairs (x floor, x ceil) per (y,z): (6,5) (7,4) (2,1) (3,0).
            const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
            constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};
            float2 f[8];
            float4 pv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t rf = corner_row(c, kf[q], hmask, base);
              if (kDbg && (a.debug & 4)) {
                pv[q] = make_float4(c.ox, c.oy, c.oz, 1.f);
                f[kc[q]] = make_float2(c.oy, c.oz);
                continue;
              }
              pv[q] = __ldg(reinterpret_cast<const float4*>(table + (rf & ~1u)));
              if (!pair) f[kc[q]] = __ldg(table + corner_row(c, kc[q], hmask, base));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t rf = corner_row(c, kf[q], hmask, base);
              const bool f_first = (rf & 1u) == 0u;
              f[kf[q]] = f_first ? make_float2(pv[q].x, pv[q].y) : make_float2(pv[q].z, pv[q].w);
              if (pair) f[kc[q]] = f_first ? make_float2(pv[q].z, pv[q].w) : make_float2(pv[q].x, pv[q].y);
            }
            const float2 r = trilerp(f, c);
            enc[2 * li] = r.x;
            enc[2 * li + 1] = r.y;
          }
          finish_tile(cur, seq);
        }
      }
      if (prof) printf("fwd coarse gather warp: total %lld cycles, waiting for a free encoding buffer %lld\n", clock64() - t_total, t_wait_a);
    } else {
      // ---- fine levels: every lane hits its own table lines, so nothing coalesces; rows are staged with cp.async (rows in flight cost
      // shared memory, not registers) kDepth (tile, level) items ahead of their blend
      int issued = 0;       // items issued so far (stage = issued % kDepth)
      TilePoint prev{};     // the tile whose last levels are still in flight
      int prev_seq = -1;

      auto issue_item = [&](const TilePoint& tp, int li) {
        const int l = l0 + li;
        const LevelCell c = level_cell(tp.pos, F.scalings[l]);
        const uint32_t base = (uint32_t)l << F.log2T;
        const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
        constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};
        const uint32_t st = stage0 + (uint32_t)(issued % kDepth) * STAGE_BYTES;
        uint32_t flags = pair ? 1u : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t rf = corner_row(c, kf[q], hmask, base);
          flags |= (rf & 1u) << (1 + q);
          if (kDbg && (a.debug & 2)) continue;
          cp_async16(st + STAGE_PAIR + q * 512 + lane * 16, table + (rf & ~1u));
          cp_async8_if(!pair, st + STAGE_SINGLE + q * 256 + lane * 8, table + corner_row(c, kc[q], hmask, base));
        }
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + STAGE_META + lane * 16), "f"(c.ox), "f"(c.oy), "f"(c.oz),
                     "f"(__uint_as_float(flags))
                     : "memory");
        cp_async_commit();
        ++issued;
      };
      // blend the item issued kDepth items ago (stage = issued % kDepth, the one the next issue overwrites) -> enc[2 li ..]
      auto blend_item = [&](int li) {
        const uint32_t st = stage0 + (uint32_t)(issued % kDepth) * STAGE_BYTES;
        float ox, oy, oz, fl;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ox), "=f"(oy), "=f"(oz), "=f"(fl) : "r"(st + STAGE_META + lane * 16) : "memory");
        const uint32_t flags = __float_as_uint(fl);
        const bool pair = flags & 1u;
        constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};
        float2 f[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 pv;
          float2 sv;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pv.x), "=f"(pv.y), "=f"(pv.z), "=f"(pv.w) : "r"(st + STAGE_PAIR + q * 512 + lane * 16) : "memory");
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(sv.x), "=f"(sv.y) : "r"(st + STAGE_SINGLE + q * 256 + lane * 8) : "memory");
          const bool f_first = ((flags >> (1 + q)) & 1u) == 0u;
          f[kf[q]] = f_first ? make_float2(pv.x, pv.y) : make_float2(pv.z, pv.w);
          const float2 nb = f_first ? make_float2(pv.z, pv.w) : make_float2(pv.x, pv.y);
          f[kc[q]] = pair ? nb : sv;
        }
        LevelCell c;
        c.ox = ox;
        c.oy = oy;
        c.oz = oz;
        const float2 r = trilerp(f, c);
        enc[2 * li] = r.x;
        enc[2 * li + 1] = r.y;
      };

      for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        const int ray0 = group * G;
        const int pts = min(G, R - ray0) * S;
        const int tiles = (pts + 127) / 128;
#pragma unroll 1
        for (int tile = 0; tile < tiles; ++tile, ++seq) {
          const TilePoint cur = tile_point<kExport>(a, ray0, pts, S, tile, row, true);
#pragma unroll
          for (int li = 0; li < kLv; ++li) {
            if (issued >= kDepth) {
              long long t0 = 0;
              if (prof) t0 = clock64();
              cp_async_wait<kDepth - 1>();
              if (prof) t_wait_b += clock64() - t0;
              constexpr int kBack = kLv - kDepth;  // the retired item is level (li + kBack) % kLv: of this tile if li >= kDepth, else of the previous
              blend_item((li + kBack) % kLv);
              if ((li + kBack) % kLv == kLv - 1) finish_tile(prev, prev_seq);  // the last level retires at li == kDepth - 1 of the NEXT tile
            }
            issue_item(cur, li);
          }
          prev = cur;
          prev_seq = seq;
        }
      }
      // drain: the last kDepth items (all of the last tile when kDepth == kLv)
      if (prev_seq >= 0) {
        cp_async_wait<0>();
#pragma unroll
        for (int li = kLv - kDepth; li < kLv; ++li) {
          blend_item(li);
          ++issued;  // walk the stages in issue order
        }
        finish_tile(prev, prev_seq);
      }
      if (prof) printf("fwd fine gather warp: total %lld cycles, waiting for a free encoding buffer %lld, for cp.async data %lld\n", clock64() - t_total, t_wait_a, t_wait_b);
    }
  } else {
    // ========================= chain warps =========================
    reg_inc<kChainRegs>();
    const int slot = warp >> 3;
    const int half = (warp >> 2) & 1;  // accumulator columns 32 half .. 32 half + 31
    const uint32_t ts = tmem0 + slot * kSlotCols;        // MMA addresses (lane 0)
    const uint32_t tr = ts + lane_off;                   // this thread's lane quarter
    const uint32_t wBase = smem_u32(smem);
    uint64_t* bar = &s_bar[slot];
    uint32_t phase = 0;
    const bool issue_warp = (warp & 7) == 0;
    const int bar_id = BAR_SLOT0 + slot;

#define FNR_SLOT_ISSUE(...)              \
  if (prof) t_mark = clock64();          \
  tmem_st_wait();                        \
  fence_before_sync();                   \
  named_bar_sync(bar_id, kSlotThreads);  \
  if (issue_warp) {                      \
    if (elect_one_sync()) {              \
      fence_after_sync();                \
      __VA_ARGS__;                       \
      mma_commit(bar);                   \
    }                                    \
    __syncwarp();                        \
  }                                      \
  if (prof) t_wait_c += clock64() - t_mark;
#define FNR_SLOT_WAIT()                    \
  if (prof) t_mark = clock64();            \
  mbar_wait(bar, phase);                   \
  phase ^= 1;                              \
  fence_after_sync();                      \
  if (prof) t_wait_d += clock64() - t_mark;
    long long t_mark = 0;

    int seq = 0;
    for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
      const int ray0 = group * G;
      const int rays_here = min(G, R - ray0);
      const int pts = rays_here * S;
      const int tiles = (pts + 127) / 128;
#pragma unroll 1
      for (int tile = 0; tile < tiles; ++tile, ++seq) {
        if ((seq & 1) != slot) continue;
        const TilePoint tp = tile_point<kExport>(a, ray0, pts, S, tile, row, false);
        const int e = seq % kEnc, use = seq / kEnc;
        {
          long long t0 = 0;
          if (prof) t0 = clock64();
          mbar_wait_sleep(&s_full[e], (uint32_t)(use & 1));  // the gather warps have written this tile's encoding + selectors
          if (prof) t_wait_a += clock64() - t0;
        }
        const bool sel = s_sel[e * 128 + row] != 0;
        if (kDbg && (a.debug & 8)) {
          named_bar_sync(bar_id, kSlotThreads);
          if (issue_warp && lane == 0) mbar_arrive(&s_empty[e]);
          if (half == 0 && tp.valid) {
            float* q = s_samples + 5 * tp.local;
            q[0] = sel ? 1.f : 0.f; q[1] = q[2] = q[3] = 0.5f; q[4] = 0.f;
          }
          continue;
        }
        // ---- base0 (A = encoding ring buffer); its completion also frees the buffer for the gather warps
        if (issue_warp) {
          if (elect_one_sync()) {
            fence_after_sync();
            issue_gemm_tsi<K_BASE0, N_BASE0>(ts + C_X, tmem0 + C_ENC + 32 * e, wBase + OFF_W_BASE0);
            mma_commit(bar);
            mma_commit(&s_empty[e]);
          }
          __syncwarp();
        }
        // colour-input k-steps that do not depend on the MLPs: 0 = sh (half 0), 1..2 = appearance embedding (half 1)
        if (half == 0) {
          const float* d = kExport ? a.E.normal : a.Rr.directions + 3 * (size_t)tp.ray;
          float sh[SHD];
          sh_degree4(d[0], d[1], d[2], sh);
          st_kstep(tr + C_C, 0, sh);
        } else {
          // both pointers stay VALID in every mode: the compiler turns the branch below into loads + selects (read-only loads are
          // speculated), so a null embedding / camera-index pointer would fault in the mean / zeros modes
          const bool per_cam = !kExport && F.appearance_mode == FNR_APP_PER_CAMERA;
          const int32_t* cip = per_cam ? a.Rr.camera_indices + tp.ray : reinterpret_cast<const int32_t*>(P.app_embedding);
          const int cam_row = per_cam ? __ldg(cip) : 0;
          const float* app = P.app_embedding + (size_t)cam_row * APP;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float v[16];
            if (per_cam) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(app) + 4 * j + q);
                v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = s_bias[B_APP + 16 * j + q];
            }
            st_kstep(tr + C_C, 1 + j, v);
          }
        }

        // ---- epilogue 1: h1 = relu(base0 + b), in place in X ; base1 -> Y
        FNR_SLOT_WAIT()
#pragma unroll
        for (int j = 0; j < 2; ++j)
          epi_inplace16(tr + C_X, 2 * half + j, [&](int n, float x) { return fmaxf(x + s_bias[B_BASE0 + n], 0.f); });
        FNR_SLOT_ISSUE(issue_gemm_tsi<K_BASE1, N_BASE1>(ts + C_Y, ts + C_X, wBase + OFF_W_BASE1))

        // ---- epilogue 2 (half 0): [h0 | geo] + b -> density ; semantic input (in place in Y) and colour-input k-step 3
        FNR_SLOT_WAIT()
        float density = 0.f;
        if (half == 0) {
          uint32_t r0[16];
          tmem_ld16(tr + C_Y, r0);
          tmem_ld_wait();
          float g[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) g[q] = __uint_as_float(r0[q]) + s_bias[B_BASE1 + q];
          density = sel ? expf(g[0]) : 0.f;
          st_kstep(tr + C_Y, 0, g);
          st_kstep(tr + C_C, 3, g);
        }
        FNR_SLOT_ISSUE(issue_gemm_tsi<K_SEM0, N_SEM0>(ts + C_Z, ts + C_Y, wBase + OFF_W_SEM0);
                       issue_gemm_tsi<K_COL0, N_COL0>(ts + C_X, ts + C_C, wBase + OFF_W_COL0))

        // ---- epilogue 3: z1 = relu(sem0 + b) in place in Z ; c1 = relu(col0 + b) in place in X
        FNR_SLOT_WAIT()
#pragma unroll
        for (int j = 0; j < 2; ++j)
          epi_inplace16(tr + C_Z, 2 * half + j, [&](int n, float x) { return fmaxf(x + s_bias[B_SEM0 + n], 0.f); });
#pragma unroll
        for (int j = 0; j < 2; ++j)
          epi_inplace16(tr + C_X, 2 * half + j, [&](int n, float x) { return fmaxf(x + s_bias[B_COL0 + n], 0.f); });
        FNR_SLOT_ISSUE(issue_gemm_tsi<K_SEMH, N_SEMH>(ts + C_Y, ts + C_Z, wBase + OFF_W_SEMH);
                       issue_gemm_tsi<K_COL1, N_COL1>(ts + C_C, ts + C_X, wBase + OFF_W_COL1))

        // ---- epilogue 4: logit ; c2 = relu(col1 + b) in place in C
        FNR_SLOT_WAIT()
        float logit = 0.f;
        if (half == 0) {
          uint32_t lg[8];
          tmem_ld8(tr + C_Y, lg);
          tmem_ld_wait();
          logit = __uint_as_float(lg[0]) + s_bias[B_SEMH];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          epi_inplace16(tr + C_C, 2 * half + j, [&](int n, float x) { return fmaxf(x + s_bias[B_COL1 + n], 0.f); });
        FNR_SLOT_ISSUE(issue_gemm_tsi<K_COL2, N_COL2>(ts + C_Y, ts + C_C, wBase + OFF_W_COL2))

        // ---- epilogue 5: rgb = sigmoid(col2 + b) ; per-sample results to shared memory
        FNR_SLOT_WAIT()
        if (half == 0) {
          uint32_t c[8];
          tmem_ld8(tr + C_Y, c);
          tmem_ld_wait();
          if (tp.valid) {
            float* q = s_samples + 5 * tp.local;
            q[0] = density;
            q[1] = sigmoidf_(__uint_as_float(c[0]) + s_bias[B_COL2]);
            q[2] = sigmoidf_(__uint_as_float(c[1]) + s_bias[B_COL2 + 1]);
            q[3] = sigmoidf_(__uint_as_float(c[2]) + s_bias[B_COL2 + 2]);
            q[4] = logit;
          }
        }
        fence_before_sync();  // order this tile's TMEM reads before the next tile's MMAs
      }

      // ---- per group: per-sample outputs, then composite one ray per warp (render) or threshold + compact (export)
      {
        long long t0 = 0;
        if (prof) t0 = clock64();
        named_bar_sync(BAR_CHAIN, kChainThreads);
        if constexpr (kExport) {
          group_export<kChainThreads>(a.E, F, s_samples, ray0, pts, S);
        } else {
          group_write_samples<kChainThreads>(a.O, s_samples, ray0, pts, S);
          if (a.composite) group_composite<kChainThreads>(a.Cm, a.Rr, s_samples, ray0, rays_here, S);
        }
        named_bar_sync(BAR_CHAIN, kChainThreads);  // s_samples is rewritten by the next group
        if (prof) t_wait_b += clock64() - t0;
      }
    }
#undef FNR_SLOT_ISSUE
#undef FNR_SLOT_WAIT
    if (prof)
      printf("fwd chain slot %d: total %lld cycles, waiting for encodings %lld, group stage (incl. its barriers) %lld, operand sync + MMA issue %lld, "
             "waiting for MMA completion %lld\n", slot, clock64() - t_total, t_wait_a, t_wait_b, t_wait_c, t_wait_d);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

int pick_rays_per_group_ws(int S) {
  if (S > kMaxGroupPoints) return 0;
  int best = 0;
  double best_waste = 2.0;
  for (int g = 1; g * S <= kMaxGroupPoints; ++g) {
    const int pts = g * S;
    const int rounds = (pts + kSlots * 128 - 1) / (kSlots * 128);
    const double waste = 1.0 - (double)pts / (rounds * kSlots * 128);
    if (waste < best_waste - 1e-9 || (waste < best_waste + 1e-9 && g > best)) {
      best_waste = waste;
      best = g;
    }
  }
  return best;
}

template <bool kExport, bool kDbg>
int configure_ws() {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_render_forward_ws_kernel<kExport, kDbg>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_render_forward_ws_kernel)");
    configured = true;
  }
  return FNR_OK;
}

}  // namespace

bool tc_ws_supported(int S) { return S >= 1 && S <= kMaxGroupPoints; }

int launch_tc_render_forward_ws(const KField& F, const KParams& P, const KRays& Rr, const KFieldOut& O, const KComposite& Cm,
                                cudaStream_t st) {
  if (Rr.R == 0) return FNR_OK;
  const int debug = getenv("FNR_DEBUG_FWD") ? atoi(getenv("FNR_DEBUG_FWD")) : 0;
  if (int rc = debug ? configure_ws<false, true>() : configure_ws<false, false>()) return rc;
  WsArgs a;
  memset(&a.E, 0, sizeof(a.E));
  a.F = F;
  a.P = P;
  a.Rr = Rr;
  a.O = O;
  a.Cm = Cm;
  a.rays_per_group = pick_rays_per_group_ws(Rr.S);
  a.debug = debug;
  a.composite = Cm.rgb || Cm.accumulation || Cm.depth || Cm.depth_index || Cm.semantics || Cm.weights;
  const int groups = (Rr.R + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  if (debug) tc_render_forward_ws_kernel<false, true><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  else tc_render_forward_ws_kernel<false, false><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_ws_kernel");
}

int launch_tc_export_ws(const KField& F, const KParams& P, const KExport& E, cudaStream_t st) {
  if (E.B == 0) return FNR_OK;
  if (int rc = configure_ws<true, false>()) return rc;
  WsArgs a;
  memset(&a, 0, sizeof(a));
  a.F = F;
  a.F.position_mode = FNR_POS_AABB;
  a.F.appearance_mode = FNR_APP_MEAN;
  a.P = P;
  a.E = E;
  a.rays_per_group = pick_rays_per_group_ws(E.S);
  const int groups = (E.B + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  tc_render_forward_ws_kernel<true, false><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_ws_kernel<export>");
}

}  // namespace fnr

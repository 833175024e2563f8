// Fused render forward for sm_100a: hash-grid gather -> base / semantic / colour MLPs on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM) -> per-ray alpha compositing, one persistent CTA
// per SM.  (fruit_nerf/fruit_field.py:168-301 + fruit_nerf/fruit_nerf.py:320-348 in one pass.)
//
// Work decomposition
//   * A CTA owns a group of G whole rays (G*S <= kMaxGroupPoints) at a time; its points are processed
//     in rounds of kSlots*128 points.  A "slot" is 8 warps = 256 threads serving the 128 rows (TMEM lanes)
//     of an M=128 MMA tile: the thread pair (r, r+128) owns point r from the gather to the heads, one
//     thread taking hash levels 0-7 / accumulator columns 0-31, the other levels 8-15 / columns 32-63
//     (two warps may address the same TMEM lane quarter), which doubles the loads in flight per point.
//     Slots only meet at the per-group compositing step, so one slot's gathers overlap the other
//     slot's tensor-core round trips.
//   * Gather: per level 8 float2 loads (read-only path), trilinear blend in registers, 32 features.
//   * MLPs: every layer is D[128,N] (TMEM, fp32) = A[128,K] (smem, bf16) x W[N,K]^T (smem, bf16), with
//     both operands split x = hi + lo (two bf16 each) and three MMAs per K-step
//     (A_hi W_hi + A_lo W_hi + A_hi W_lo): ~2^-16 relative error per product, i.e. fp32-class parity
//     (north_star 1e-3) at 3x the tensor work -- affordable because the path is gather-bound
//     (DESIGN.md).  The epilogue thread reads its row with tcgen05.ld, adds the bias, applies ReLU,
//     re-splits and stores the next layer's A row (canonical K-major layout, 16-byte vector stores).
//     The last semantic layer (no activation) is folded into the 1-logit head at weight-staging time.
//   * Compositing: per-sample density / rgb / logit stay in shared memory; one warp per ray scans.
//
// Shapes: the fruit_nerf family (geo 15, semantic 15-64-64, colour 63-64-64-3).  The fruit_nerf_big family is dispatched to
// fnr_tc_big.cu (activations in tensor memory); everything else is served by the simt kernels (tc_supported() == false).
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
constexpr int kSlotThreads = 256;  // 2 threads per point (column / level halves)
constexpr int kCtaThreads = kSlots * kSlotThreads;
constexpr int kMaxGroupPoints = 768;
constexpr int kTmemColsPerSlot = 160;


// ---- small-family dimensions -------------------------------------------------------------------
constexpr int GEO = 15, ENC = 32, H = 64, APP = 32, SHD = 16;
// padded GEMM shapes (K multiple of 16, N multiple of 16)
constexpr int K_BASE0 = 32, N_BASE0 = 64;
constexpr int K_BASE1 = 64, N_BASE1 = 16;
constexpr int K_SEM0 = 16, N_SEM0 = 64;
constexpr int K_SEMH = 64, N_SEMH = 16;   // folded (semantic layer 1) x head
constexpr int K_COL0 = 64, N_COL0 = 64;   // K order: [sh 16 | app 32 | geo 15 | 0]
constexpr int K_COL1 = 64, N_COL1 = 64;
constexpr int K_COL2 = 64, N_COL2 = 16;

// shared-memory map (bytes).  Weight tiles: hi then lo, canonical layout with ROWS = N.
constexpr int OFF_W_BASE0 = 0;
constexpr int OFF_W_BASE1 = OFF_W_BASE0 + 2 * wbytes(N_BASE0, K_BASE0);
constexpr int OFF_W_SEM0 = OFF_W_BASE1 + 2 * wbytes(N_BASE1, K_BASE1);
constexpr int OFF_W_SEMH = OFF_W_SEM0 + 2 * wbytes(N_SEM0, K_SEM0);
constexpr int OFF_W_COL0 = OFF_W_SEMH + 2 * wbytes(N_SEMH, K_SEMH);
constexpr int OFF_W_COL1 = OFF_W_COL0 + 2 * wbytes(N_COL0, K_COL0);
constexpr int OFF_W_COL2 = OFF_W_COL1 + 2 * wbytes(N_COL1, K_COL1);
constexpr int OFF_BIAS = OFF_W_COL2 + 2 * wbytes(N_COL2, K_COL2);
// biases (floats): base0[64] base1[16] sem0[64] semh[16] col0[64] col1[64] col2[16] app[32]
constexpr int B_BASE0 = 0, B_BASE1 = 64, B_SEM0 = 80, B_SEMH = 144, B_COL0 = 160, B_COL1 = 224, B_COL2 = 288, B_APP = 304,
              B_COUNT = 336;
constexpr int OFF_TILES = OFF_BIAS + B_COUNT * 4;
// per-slot activation tiles: P (K=64 hi+lo), Q (K=64 hi+lo; the K=32 encoding tile aliases it), S (K=16 hi+lo)
constexpr int TILE_P = 0, TILE_Q = 2 * 128 * 64 * 2, TILE_S = 2 * TILE_Q, SLOT_TILE_BYTES = TILE_S + 2 * 128 * 16 * 2;
constexpr int OFF_SAMPLES = OFF_TILES + kSlots * SLOT_TILE_BYTES;  // [kMaxGroupPoints][5] floats
constexpr int OFF_END = OFF_SAMPLES + kMaxGroupPoints * 5 * 4;
constexpr int kSmemBytes = OFF_END + 1024;  // + alignment slack
static_assert(kSmemBytes <= 227 * 1024, "shared-memory budget");
static_assert(OFF_TILES % 16 == 0 && SLOT_TILE_BYTES % 16 == 0 && OFF_BIAS % 16 == 0, "alignment");

// TMEM column map inside a slot's kTmemColsPerSlot window
constexpr int C_R0 = 0;    // 64 cols: base0 out, later colour0 out, colour1 out
constexpr int C_R1 = 64;   // 16 cols: base1 out, later semantic-head out, colour2 out
constexpr int C_R2 = 80;   // 64 cols: semantic0 out

struct TcArgs {
  KField F;
  KParams P;
  KRays Rr;
  KFieldOut O;
  KComposite Cm;
  int rays_per_group;
  int composite;
  int debug;  // FNR_DEBUG_FWD: block 0 prints per-phase cycle totals (timing experiments only)
  KExport E;  // kExport instantiation only: rays come from (origins, normal, bins) and the per-group stage compacts
};

// 32-column epilogue of the thread pair (row, half): v = f(column, accumulator) -> chunks 4*half.. of a K=64 tile.
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

template <bool kExport>
__global__ void __launch_bounds__(kCtaThreads, 1) tc_render_forward_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_bar[kSlots];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = warp >> 3;              // 8 warps per slot
  const int half = (warp >> 2) & 1;        // which 32 of 64 accumulator columns / which 8 of 16 levels
  const int row = (warp & 3) * 32 + lane;  // TMEM lane = point of this thread pair (warp w may touch lanes 32*(w%4)..)
  const KParams& P = a.P;
  const KField& F = a.F;
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* s_samples = reinterpret_cast<float*>(smem + OFF_SAMPLES);

  // ---- one-time setup: TMEM, barriers, weights ------------------------------------------------
  if (warp == 0) tmem_alloc(&s_tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&s_bar[i], 1);
    mbar_fence_init();
  }
  stage_weight<kCtaThreads, N_BASE0, K_BASE0>(smem + OFF_W_BASE0, [&](int n, int k) { return __ldg(P.base_w[0] + n * ENC + k); });
  stage_weight<kCtaThreads, N_BASE1, K_BASE1>(smem + OFF_W_BASE1, [&](int n, int k) { return __ldg(P.base_w[1] + n * H + k); });
  stage_weight<kCtaThreads, N_SEM0, K_SEM0>(smem + OFF_W_SEM0, [&](int n, int k) { return k < GEO ? __ldg(P.sem_w[0] + n * GEO + k) : 0.f; });
  // fold: logit = head_w . (W1 z + b1) + head_b  =>  row 0 of the N=16 tile is head_w^T W1
  stage_weight<kCtaThreads, N_SEMH, K_SEMH>(smem + OFF_W_SEMH, [&](int n, int k) {
    if (n != 0) return 0.f;
    float acc = 0.f;
    for (int j = 0; j < H; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_w[1] + j * H + k), acc);
    return acc;
  });
  // colour layer 0 with the K order [sh | app | geo | 0] (torch order is [sh | geo | app])
  stage_weight<kCtaThreads, N_COL0, K_COL0>(smem + OFF_W_COL0, [&](int n, int k) {
    const float* w = P.col_w[0] + n * (SHD + GEO + APP);
    if (k < SHD) return __ldg(w + k);
    if (k < SHD + APP) return __ldg(w + SHD + GEO + (k - SHD));
    if (k < SHD + APP + GEO) return __ldg(w + SHD + (k - SHD - APP));
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

  const uint32_t tmem_slot = s_tmem_base + slot * kTmemColsPerSlot;
  const uint32_t tmem_row = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* tiles = smem + OFF_TILES + slot * SLOT_TILE_BYTES;
  uint8_t* tP = tiles + TILE_P;
  uint8_t* tQ = tiles + TILE_Q;
  uint8_t* tS = tiles + TILE_S;
  const uint32_t aP = smem_u32(tP), aQ = smem_u32(tQ), aS = smem_u32(tS);
  const uint32_t wBase = smem_u32(smem);
  uint64_t* bar = &s_bar[slot];
  uint32_t phase = 0;
  const bool issue_warp = (warp & 7) == 0;
  const int bar_id = 1 + slot;

  const int S = kExport ? a.E.S : a.Rr.S, R = kExport ? a.E.B : a.Rr.R;
  const int G = a.rays_per_group;
  const int num_groups = (R + G - 1) / G;
  const float2* __restrict__ table = reinterpret_cast<const float2*>(P.hash_table);
  const uint32_t hmask = (1u << F.log2T) - 1u;

#define FNR_SLOT_ISSUE(...)              \
  fence_async_smem();                    \
  fence_before_sync();                   \
  named_bar_sync(bar_id, kSlotThreads);  \
  if (issue_warp) {                      \
    if (elect_one_sync()) {              \
      fence_after_sync();                \
      __VA_ARGS__;                       \
      mma_commit(bar);                   \
    }                                    \
    __syncwarp();                        \
  }
#define FNR_SLOT_WAIT()  \
  mbar_wait(bar, phase); \
  phase ^= 1;            \
  fence_after_sync();

  long long t_gather = 0, t_chain = 0, t_group = 0, t_mark = 0;
  const bool prof = a.debug && blockIdx.x == 0 && (tid & 255) == 0;
#define FNR_TICK(acc)                  \
  if (prof) {                          \
    const long long now_ = clock64();  \
    acc += now_ - t_mark;              \
    t_mark = now_;                     \
  }
  if (prof) t_mark = clock64();

  for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
    const int ray0 = group * G;
    const int rays_here = min(G, R - ray0);
    const int pts = rays_here * S;
    const int rounds = (pts + kSlots * 128 - 1) / (kSlots * 128);

#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      const int local = rd * (kSlots * 128) + slot * 128 + row;  // point index inside the group
      const bool valid = local < pts;
      const int lc = valid ? local : pts - 1;
      const int ray = ray0 + lc / S;
      const size_t gp = (size_t)ray0 * S + lc;  // global point index
      const float* o = (kExport ? a.E.origins : a.Rr.origins) + 3 * (size_t)ray;
      const float* d = kExport ? a.E.normal : a.Rr.directions + 3 * (size_t)ray;
      float t0, t1;
      if constexpr (kExport) {
        export_interval(a.E, ray, lc % S, t0, t1);
      } else {
        t0 = __ldg(a.Rr.starts + gp);
        t1 = __ldg(a.Rr.ends + gp);
      }
      bool sel;
      const Vec3 pos = field_position(o, d, t0, t1, F.position_mode, F.aabb, sel);

      // ---- gather + trilinear blend: this thread's 8 levels -> chunks 2*half, 2*half+1 of the encoding tile (aliases Q)
      {
        float enc[16];
#pragma unroll
        for (int li = 0; li < 8; ++li) {
          const int l = 8 * half + li;
          const LevelCell c = level_cell(pos, F.scalings[l]);
          const uint32_t base = (uint32_t)l << F.log2T;
          // x-neighbours (floor x even, ceil = floor + 1) are adjacent table rows r, r^1: one 16-byte load
          // fetches both; only lanes with an odd floor x issue the second, 8-byte load.  Corner pairs
          // (x = floor, x = ceil) per (y,z): (6,5) (7,4) (2,1) (3,0).  Fewer LDG instructions and fewer
          // distinct lines per instruction = fewer L1TEX wavefront replays (the gather's bound).
          const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);
          constexpr int kf[4] = {6, 7, 2, 3}, kc[4] = {5, 4, 1, 0};
          float2 f[8];
          float4 pv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t rf = corner_row(c, kf[q], hmask, base);
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
        if (a.O.stash_encoding && valid) {
          float* st = a.O.stash_encoding + gp * ENC + 16 * half;  // 64-byte aligned: two whole sectors per thread
          st_global_v8(st, enc[0], enc[1], enc[2], enc[3], enc[4], enc[5], enc[6], enc[7]);
          st_global_v8(st + 8, enc[8], enc[9], enc[10], enc[11], enc[12], enc[13], enc[14], enc[15]);
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = enc[8 * jj + q];
          store_chunk(tQ, 128 * K_BASE0 * 2, row, 2 * half + jj, v);
        }
      }
      FNR_TICK(t_gather)
      FNR_SLOT_ISSUE(issue_gemm<K_BASE0, N_BASE0>(tmem_slot + C_R0, aQ, wBase + OFF_W_BASE0))

      // ---- epilogue 1: h1 = relu(base0 + b) -> P ; then the geo-independent colour-input chunks -> Q ------
      FNR_SLOT_WAIT()
      epi32(tmem_row + C_R0 + 32 * half, tP, row, half, [&](int n, float x) { return fmaxf(x + s_bias[B_BASE0 + n], 0.f); });
      FNR_SLOT_ISSUE(issue_gemm<K_BASE1, N_BASE1>(tmem_slot + C_R1, aP, wBase + OFF_W_BASE1))
      // colour input K order [sh 0..15 | app 0..31 | geo | 0]: chunks 0,1 = sh (half 0), 2..5 = app (half 1).
      // (the encoding tile in Q is dead: its GEMM completed before epilogue 1 ran)
      if (half == 0) {
        float sh[SHD];
        sh_degree4(d[0], d[1], d[2], sh);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sh[8 * j + q];
          store_chunk(tQ, 128 * 64 * 2, row, j, v);
        }
      } else {
        const float* app = (!kExport && F.appearance_mode == FNR_APP_PER_CAMERA)
                               ? P.app_embedding + (size_t)__ldg(a.Rr.camera_indices + ray) * APP
                               : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8];
          if (app) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(app) + 2 * j), w = __ldg(reinterpret_cast<const float4*>(app) + 2 * j + 1);
            v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; v[4] = w.x; v[5] = w.y; v[6] = w.z; v[7] = w.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = s_bias[B_APP + 8 * j + q];
          }
          store_chunk(tQ, 128 * 64 * 2, row, 2 + j, v);
        }
      }

      // ---- epilogue 2: [h0 | geo] ; density ; geo -> S (semantic input) and chunks 6,7 of Q --------
      FNR_SLOT_WAIT()
      float density = 0.f;
      {
        uint32_t r0[16];
        tmem_ld16(tmem_row + C_R1, r0);
        tmem_ld_wait();
        float g[8];
        if (half == 0) {
          density = sel ? expf(__uint_as_float(r0[0]) + s_bias[B_BASE1]) : 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = __uint_as_float(r0[1 + q]) + s_bias[B_BASE1 + 1 + q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = q < 7 ? __uint_as_float(r0[9 + q]) + s_bias[B_BASE1 + 9 + q] : 0.f;
        }
        store_chunk(tS, 128 * 16 * 2, row, half, g);
        store_chunk(tQ, 128 * 64 * 2, row, 6 + half, g);
      }
      FNR_SLOT_ISSUE(issue_gemm<K_SEM0, N_SEM0>(tmem_slot + C_R2, aS, wBase + OFF_W_SEM0);
                     issue_gemm<K_COL0, N_COL0>(tmem_slot + C_R0, aQ, wBase + OFF_W_COL0))

      // ---- epilogue 3: z1 = relu(sem0 + b) -> P ; c1 = relu(col0 + b) -> Q -------------------------
      FNR_SLOT_WAIT()
      epi32(tmem_row + C_R2 + 32 * half, tP, row, half, [&](int n, float x) { return fmaxf(x + s_bias[B_SEM0 + n], 0.f); });
      epi32(tmem_row + C_R0 + 32 * half, tQ, row, half, [&](int n, float x) { return fmaxf(x + s_bias[B_COL0 + n], 0.f); });
      FNR_SLOT_ISSUE(issue_gemm<K_SEMH, N_SEMH>(tmem_slot + C_R1, aP, wBase + OFF_W_SEMH);
                     issue_gemm<K_COL1, N_COL1>(tmem_slot + C_R0, aQ, wBase + OFF_W_COL1))

      // ---- epilogue 4: logit ; c2 = relu(col1 + b) -> P --------------------------------------------
      FNR_SLOT_WAIT()
      float logit = 0.f;
      if (half == 0) {
        uint32_t lg[8];
        tmem_ld8(tmem_row + C_R1, lg);
        tmem_ld_wait();
        logit = __uint_as_float(lg[0]) + s_bias[B_SEMH];
      }
      epi32(tmem_row + C_R0 + 32 * half, tP, row, half, [&](int n, float x) { return fmaxf(x + s_bias[B_COL1 + n], 0.f); });
      FNR_SLOT_ISSUE(issue_gemm<K_COL2, N_COL2>(tmem_slot + C_R1, aP, wBase + OFF_W_COL2))

      // ---- epilogue 5: rgb = sigmoid(col2 + b) ; per-sample results to shared memory ---------------
      FNR_SLOT_WAIT()
      if (half == 0) {
        uint32_t c[8];
        tmem_ld8(tmem_row + C_R1, c);
        tmem_ld_wait();
        if (valid) {
          float* q = s_samples + 5 * local;
          q[0] = density;
          q[1] = sigmoidf_(__uint_as_float(c[0]) + s_bias[B_COL2]);
          q[2] = sigmoidf_(__uint_as_float(c[1]) + s_bias[B_COL2 + 1]);
          q[3] = sigmoidf_(__uint_as_float(c[2]) + s_bias[B_COL2 + 2]);
          q[4] = logit;
        }
      }
      fence_before_sync();  // order this round's TMEM reads before the next round's MMAs
      FNR_TICK(t_chain)
    }
#undef FNR_SLOT_ISSUE
#undef FNR_SLOT_WAIT

    // ---- per-group: write per-sample outputs, then composite one ray per warp (render) or threshold + compact (export)
    __syncthreads();
    if constexpr (kExport) {
      group_export<kCtaThreads>(a.E, F, s_samples, ray0, pts, S);
    } else {
      group_write_samples<kCtaThreads>(a.O, s_samples, ray0, pts, S);
      if (a.composite) group_composite<kCtaThreads>(a.Cm, a.Rr, s_samples, ray0, rays_here, S);
    }
    __syncthreads();  // s_samples is rewritten by the next group
    FNR_TICK(t_group)
  }
  if (prof) printf("fwd slot %d: gather %lld chain %lld group-stage %lld cycles\n", slot, t_gather, t_chain, t_group);
#undef FNR_TICK

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

int pick_rays_per_group(int S) {
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

}  // namespace

bool tc_supported(Family fam, const KField& F, const KRays& Rr) {
  (void)F;
  if (fam == kFamilyBig) return tc_big_supported(Rr.S);
  return fam == kFamilySmall && Rr.S >= 1 && Rr.S <= kMaxGroupPoints;
}

bool tc_export_supported(Family fam, const KExport& E) {
  if (fam == kFamilyBig) return tc_big_supported(E.S);
  return fam == kFamilySmall && E.S >= 1 && E.S <= kMaxGroupPoints;
}

// FNR_FWD_V1=1 selects the round-1 kernel of this file (gather and chain in the same warps) for A/B timing
static bool use_v1_forward() {
  static int v = -1;
  if (v < 0) v = getenv("FNR_FWD_V1") != nullptr;
  return v != 0;
}

template <bool kExport>
static int configure_tc_forward() {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_render_forward_kernel<kExport>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_render_forward_kernel)");
    configured = true;
  }
  return FNR_OK;
}

int launch_tc_render_forward(Family fam, const KField& F, const KParams& P, const KRays& Rr, const KFieldOut& O,
                             const KComposite& Cm, cudaStream_t st) {
  if (!tc_supported(fam, F, Rr)) {
    set_error("tcgen05 render kernel does not support this shape");
    return FNR_ERR_UNSUPPORTED;
  }
  if (Rr.R == 0) return FNR_OK;
  if (fam == kFamilyBig) return launch_tc_render_forward_big(F, P, Rr, O, Cm, st);
  if (!use_v1_forward()) return launch_tc_render_forward_ws(F, P, Rr, O, Cm, st);
  if (int rc = configure_tc_forward<false>()) return rc;
  TcArgs a;
  memset(&a.E, 0, sizeof(a.E));
  a.F = F;
  a.P = P;
  a.Rr = Rr;
  a.O = O;
  a.Cm = Cm;
  a.rays_per_group = pick_rays_per_group(Rr.S);
  a.debug = getenv("FNR_DEBUG_FWD") != nullptr;
  a.composite = Cm.rgb || Cm.accumulation || Cm.depth || Cm.depth_index || Cm.semantics || Cm.weights;
  const int groups = (Rr.R + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  tc_render_forward_kernel<false><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_kernel");
}

// The export path reuses the fused forward (AABB positions, mean appearance embedding) and replaces the
// compositing stage by the three threshold selections + stream compaction.
int launch_tc_export(Family fam, const KField& F, const KParams& P, const KExport& E, cudaStream_t st) {
  if (!tc_export_supported(fam, E)) {
    set_error("tcgen05 export kernel does not support this shape");
    return FNR_ERR_UNSUPPORTED;
  }
  if (E.B == 0) return FNR_OK;
  if (fam == kFamilyBig) return launch_tc_export_big(F, P, E, st);
  if (!use_v1_forward()) return launch_tc_export_ws(F, P, E, st);
  if (int rc = configure_tc_forward<true>()) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.F = F;
  a.F.position_mode = FNR_POS_AABB;
  a.F.appearance_mode = FNR_APP_MEAN;
  a.P = P;
  a.E = E;
  a.rays_per_group = pick_rays_per_group(E.S);
  const int groups = (E.B + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  tc_render_forward_kernel<true><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_kernel<export>");
}

}  // namespace fnr

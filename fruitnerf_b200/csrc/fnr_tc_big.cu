// Fused render / export forward of the fruit_nerf_big family (geo 30, semantic 30-128-128-64-1, colour 78-64-64-3)
// on the 5th-gen tensor cores.  Same decomposition as fnr_tc.cu (thread pair (row, half) per point, bf16 hi/lo split
// operands with three MMAs per K step, per-group compositing in shared memory) with two differences forced by the
// size of the network:
//   * the 144 KB of split weight tiles fill shared memory, so the ACTIVATIONS never go there: every epilogue reads
//     its accumulator row from tensor memory (tcgen05.ld), applies bias / ReLU, re-splits and writes the next layer's
//     A operand straight back into tensor memory (tcgen05.st); the MMAs take A from TMEM (tcgen05.mma [d], [a], b-desc).
//   * one 128-row slot per CTA (the A tiles + accumulators of the 128-wide semantic layers need 384 of the 512 columns).
// TMEM plan (columns): A0 = [0,128), A1 = [128,256) operand regions, D0 = [256,384), D1 = [384,512) accumulators.
//   An A tile of depth K sits at the start of its region: hi halves in columns [0, K/2), lo halves in [K/2, K)
//   (two bf16 per 32-bit column, lane = row); K-step ks addresses hi at +8 ks and lo at K/2 + 8 ks.
// Round trips per 128 points: base0 -> base1 -> [semantic0 || colour0] -> [semantic1 || colour1] -> [folded head || colour2].
// base1's 32 outputs [h0 | geo 0..29 | pad] are fed as they are to the semantic / colour inputs (the weight rows of the
// h0 and pad positions are zero), so no thread ever needs another thread's accumulator column.
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

constexpr int kCtaThreads = 256;
constexpr int kMaxGroupPoints = 768;

constexpr int GEO = 30, ENC = 32, H = 64, APP = 32, SHD = 16, SW = 128, SOUT = 64;
constexpr int K_BASE0 = 32, N_BASE0 = 64;
constexpr int K_BASE1 = 64, N_BASE1 = 32;   // outputs [h0 | geo | pad]
constexpr int K_SEM0 = 32, N_SEM0 = 128;    // K order [h0-slot (zero row) | geo | pad]
constexpr int K_SEM1 = 128, N_SEM1 = 128;
constexpr int K_SEMF = 128, N_SEMF = 16;    // folded: head_w . (W_sem2 z + b_sem2) + head_b -> row 0
constexpr int K_COL0 = 80, N_COL0 = 64;     // K order [sh 16 | app 32 | h0-slot, geo, pad]
constexpr int K_COL1 = 64, N_COL1 = 64;
constexpr int K_COL2 = 64, N_COL2 = 16;

constexpr int OFF_W_BASE0 = 0;
constexpr int OFF_W_BASE1 = OFF_W_BASE0 + 2 * wbytes(N_BASE0, K_BASE0);
constexpr int OFF_W_SEM0 = OFF_W_BASE1 + 2 * wbytes(N_BASE1, K_BASE1);
constexpr int OFF_W_SEM1 = OFF_W_SEM0 + 2 * wbytes(N_SEM0, K_SEM0);
constexpr int OFF_W_SEMF = OFF_W_SEM1 + 2 * wbytes(N_SEM1, K_SEM1);
constexpr int OFF_W_COL0 = OFF_W_SEMF + 2 * wbytes(N_SEMF, K_SEMF);
constexpr int OFF_W_COL1 = OFF_W_COL0 + 2 * wbytes(N_COL0, K_COL0);
constexpr int OFF_W_COL2 = OFF_W_COL1 + 2 * wbytes(N_COL1, K_COL1);
constexpr int OFF_BIAS = OFF_W_COL2 + 2 * wbytes(N_COL2, K_COL2);
constexpr int B_BASE0 = 0, B_BASE1 = 64, B_SEM0 = 96, B_SEM1 = 224, B_SEMF = 352, B_COL0 = 368, B_COL1 = 432, B_COL2 = 496, B_APP = 512,
              B_COUNT = 544;
constexpr int OFF_SAMPLES = OFF_BIAS + B_COUNT * 4;
constexpr int OFF_END = OFF_SAMPLES + kMaxGroupPoints * 5 * 4;
constexpr int kSmemBytes = OFF_END + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared-memory budget");
static_assert(OFF_BIAS % 16 == 0 && OFF_SAMPLES % 16 == 0, "alignment");

constexpr int R_A0 = 0, R_A1 = 128, R_D0 = 256, R_D1 = 384;

struct BigArgs {
  KField F;
  KParams P;
  KRays Rr;
  KFieldOut O;
  KComposite Cm;
  int rays_per_group;
  int composite;
  KExport E;
};

// 16 consecutive K elements [k0, k0+16) of this thread's row -> A tile (depth K) at region base `ab` (lane-adjusted address)
template <int K>
__device__ __forceinline__ void st_a16(uint32_t ab, int k0, const float (&v)[16]) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_pack_bf16x2(v[2 * q], v[2 * q + 1], h[q], l[q]);
  tmem_st8(ab + (k0 >> 1), h);
  tmem_st8(ab + K / 2 + (k0 >> 1), l);
}

// Epilogue of NC accumulator columns starting at column c0 of the accumulator at `acc` (lane-adjusted): v = f(n, x) -> A
// tile of depth K at `ab`, K positions kdst0 + (n - c0).
template <int NC, int K, class Fn>
__device__ __forceinline__ void epi_ts(uint32_t acc, int c0, uint32_t ab, int kdst0, Fn f) {
#pragma unroll
  for (int j = 0; j < NC / 16; ++j) {
    uint32_t r[16];
    tmem_ld16(acc + c0 + 16 * j, r);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = f(c0 + 16 * j + q, __uint_as_float(r[q]));
    st_a16<K>(ab, kdst0 + 16 * j, v);
  }
}

// D[128,N] = A[128,K] (tensor memory, hi/lo) x W[N,K]^T (shared memory, hi/lo): three MMAs per K step, one thread.
template <int K, int N>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_hi) {
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

template <bool kExport>
__global__ void __launch_bounds__(kCtaThreads, 1) tc_render_forward_big_kernel(const __grid_constant__ BigArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = (warp >> 2) & 1;
  const int row = (warp & 3) * 32 + lane;
  const KParams& P = a.P;
  const KField& F = a.F;
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* s_samples = reinterpret_cast<float*>(smem + OFF_SAMPLES);

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
  stage_weight<kCtaThreads, N_SEMF, K_SEMF>(smem + OFF_W_SEMF, [&](int n, int k) {
    if (n != 0) return 0.f;
    float acc = 0.f;
    for (int j = 0; j < SOUT; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_w[2] + j * SW + k), acc);
    return acc;
  });
  stage_weight<kCtaThreads, N_COL0, K_COL0>(smem + OFF_W_COL0, [&](int n, int k) {
    const float* w = P.col_w[0] + n * (SHD + GEO + APP);  // torch order [sh | geo | app]
    if (k < SHD) return __ldg(w + k);
    if (k < SHD + APP) return __ldg(w + SHD + GEO + (k - SHD));
    const int g = k - SHD - APP - 1;  // position 0 of the last block is the h0 slot
    return (g >= 0 && g < GEO) ? __ldg(w + SHD + g) : 0.f;
  });
  stage_weight<kCtaThreads, N_COL1, K_COL1>(smem + OFF_W_COL1, [&](int n, int k) { return __ldg(P.col_w[1] + n * H + k); });
  stage_weight<kCtaThreads, N_COL2, K_COL2>(smem + OFF_W_COL2, [&](int n, int k) { return n < 3 ? __ldg(P.col_w[2] + n * H + k) : 0.f; });
  for (int i = tid; i < B_COUNT; i += kCtaThreads) {
    float v = 0.f;
    if (i < B_BASE1) v = __ldg(P.base_b[0] + i);
    else if (i < B_SEM0) v = (i - B_BASE1) < 1 + GEO ? __ldg(P.base_b[1] + (i - B_BASE1)) : 0.f;
    else if (i < B_SEM1) v = __ldg(P.sem_b[0] + (i - B_SEM0));
    else if (i < B_SEMF) v = __ldg(P.sem_b[1] + (i - B_SEM1));
    else if (i == B_SEMF) {
      float acc = __ldg(P.head_b);
      for (int j = 0; j < SOUT; ++j) acc = fmaf(__ldg(P.head_w + j), __ldg(P.sem_b[2] + j), acc);
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

  const uint32_t tb = s_tmem_base;                                  // MMA operand / accumulator addresses (lane 0)
  const uint32_t tr = tb + ((uint32_t)((warp & 3) * 32) << 16);     // this thread's lane quarter
  const uint32_t wBase = smem_u32(smem);
  uint32_t phase = 0;

  const int S = kExport ? a.E.S : a.Rr.S, R = kExport ? a.E.B : a.Rr.R;
  const int G = a.rays_per_group;
  const int num_groups = (R + G - 1) / G;
  const float2* __restrict__ table = reinterpret_cast<const float2*>(P.hash_table);
  const uint32_t hmask = (1u << F.log2T) - 1u;

#define FNR_ISSUE(...)          \
  tmem_st_wait();               \
  fence_before_sync();          \
  __syncthreads();              \
  if (warp == 0) {              \
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

  for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
    const int ray0 = group * G;
    const int rays_here = min(G, R - ray0);
    const int pts = rays_here * S;
    const int rounds = (pts + 127) / 128;

#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      const int local = rd * 128 + row;
      const bool valid = local < pts;
      const int lc = valid ? local : pts - 1;
      const int ray = ray0 + lc / S;
      const size_t gp = (size_t)ray0 * S + lc;
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

      // ---- gather: this thread's 8 levels -> K elements [16 half, 16 half + 16) of the encoding tile in A0 ----
      {
        float enc[16];
#pragma unroll
        for (int li = 0; li < 8; ++li) {
          const int l = 8 * half + li;
          const LevelCell c = level_cell(pos, F.scalings[l]);
          const uint32_t base = (uint32_t)l << F.log2T;
          const bool pair = ((c.hx[0] & 1u) == 0u) && (c.hx[1] == c.hx[0] + 1u);  // x-neighbours share a 16-byte line
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
        if (!kExport && a.O.stash_encoding && valid) {
          float* st = a.O.stash_encoding + gp * ENC + 16 * half;  // 64-byte aligned: two whole sectors per thread
          st_global_v8(st, enc[0], enc[1], enc[2], enc[3], enc[4], enc[5], enc[6], enc[7]);
          st_global_v8(st + 8, enc[8], enc[9], enc[10], enc[11], enc[12], enc[13], enc[14], enc[15]);
        }
        st_a16<K_BASE0>(tr + R_A0, 16 * half, enc);
      }
      FNR_ISSUE(issue_gemm_ts<K_BASE0, N_BASE0>(tb + R_D0, tb + R_A0, wBase + OFF_W_BASE0))

      // ---- epilogue 1: h1 = relu(base0 + b) -> A1 (K = 64); colour-input blocks that do not depend on geo -> A0 (K = 80)
      FNR_WAIT()
      epi_ts<32, K_BASE1>(tr + R_D0, 32 * half, tr + R_A1, 32 * half, [&](int n, float x) { return fmaxf(x + s_bias[B_BASE0 + n], 0.f); });
      if (half == 0) {
        float sh[16];
        sh_degree4(d[0], d[1], d[2], sh);
        st_a16<K_COL0>(tr + R_A0, 0, sh);
      } else {
        const float* app = (!kExport && F.appearance_mode == FNR_APP_PER_CAMERA)
                               ? P.app_embedding + (size_t)__ldg(a.Rr.camera_indices + ray) * APP
                               : nullptr;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = app ? __ldg(app + 16 * j + q) : s_bias[B_APP + 16 * j + q];
          st_a16<K_COL0>(tr + R_A0, SHD + 16 * j, v);
        }
      }
      FNR_ISSUE(issue_gemm_ts<K_BASE1, N_BASE1>(tb + R_D1, tb + R_A1, wBase + OFF_W_BASE1))

      // ---- epilogue 2: [h0 | geo | pad] -> semantic input (A1 + 64, K = 32) and the last block of the colour input ----
      FNR_WAIT()
      float density = 0.f;
      {
        uint32_t r0[16];
        tmem_ld16(tr + R_D1 + 16 * half, r0);
        tmem_ld_wait();
        float g[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) g[q] = __uint_as_float(r0[q]) + s_bias[B_BASE1 + 16 * half + q];
        if (half == 0) density = sel ? expf(g[0]) : 0.f;
        st_a16<K_SEM0>(tr + R_A1 + 64, 16 * half, g);
        st_a16<K_COL0>(tr + R_A0, SHD + APP + 16 * half, g);
      }
      FNR_ISSUE(issue_gemm_ts<K_SEM0, N_SEM0>(tb + R_D0, tb + R_A1 + 64, wBase + OFF_W_SEM0);
                issue_gemm_ts<K_COL0, N_COL0>(tb + R_D1 + 64, tb + R_A0, wBase + OFF_W_COL0))

      // ---- epilogue 3: z1 = relu(sem0 + b) -> A1 (K = 128) ; c1 = relu(col0 + b) -> A0 (K = 64) ----
      FNR_WAIT()
      epi_ts<64, K_SEM1>(tr + R_D0, 64 * half, tr + R_A1, 64 * half, [&](int n, float x) { return fmaxf(x + s_bias[B_SEM0 + n], 0.f); });
      epi_ts<32, K_COL1>(tr + R_D1 + 64, 32 * half, tr + R_A0, 32 * half, [&](int n, float x) { return fmaxf(x + s_bias[B_COL0 + n], 0.f); });
      FNR_ISSUE(issue_gemm_ts<K_SEM1, N_SEM1>(tb + R_D0, tb + R_A1, wBase + OFF_W_SEM1);
                issue_gemm_ts<K_COL1, N_COL1>(tb + R_D1, tb + R_A0, wBase + OFF_W_COL1))

      // ---- epilogue 4: z2 = relu(sem1 + b) -> A1 ; c2 = relu(col1 + b) -> A0 ----
      FNR_WAIT()
      epi_ts<64, K_SEMF>(tr + R_D0, 64 * half, tr + R_A1, 64 * half, [&](int n, float x) { return fmaxf(x + s_bias[B_SEM1 + n], 0.f); });
      epi_ts<32, K_COL2>(tr + R_D1, 32 * half, tr + R_A0, 32 * half, [&](int n, float x) { return fmaxf(x + s_bias[B_COL1 + n], 0.f); });
      FNR_ISSUE(issue_gemm_ts<K_SEMF, N_SEMF>(tb + R_D0, tb + R_A1, wBase + OFF_W_SEMF);
                issue_gemm_ts<K_COL2, N_COL2>(tb + R_D1 + 64, tb + R_A0, wBase + OFF_W_COL2))

      // ---- epilogue 5: logit, rgb = sigmoid(col2 + b) -> per-sample results in shared memory ----
      FNR_WAIT()
      if (half == 0) {
        uint32_t lg[8], c[8];
        tmem_ld8(tr + R_D0, lg);
        tmem_ld8(tr + R_D1 + 64, c);
        tmem_ld_wait();
        if (valid) {
          float* q = s_samples + 5 * local;
          q[0] = density;
          q[1] = sigmoidf_(__uint_as_float(c[0]) + s_bias[B_COL2]);
          q[2] = sigmoidf_(__uint_as_float(c[1]) + s_bias[B_COL2 + 1]);
          q[3] = sigmoidf_(__uint_as_float(c[2]) + s_bias[B_COL2 + 2]);
          q[4] = __uint_as_float(lg[0]) + s_bias[B_SEMF];
        }
      }
      fence_before_sync();  // this round's TMEM reads are ordered before the next round's writes / MMAs
    }
#undef FNR_ISSUE
#undef FNR_WAIT

    __syncthreads();
    if constexpr (kExport) {
      group_export<kCtaThreads>(a.E, F, s_samples, ray0, pts, S);
    } else {
      group_write_samples<kCtaThreads>(a.O, s_samples, ray0, pts, S);
      if (a.composite) group_composite<kCtaThreads>(a.Cm, a.Rr, s_samples, ray0, rays_here, S);
    }
    __syncthreads();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem_base, 512);
}

int pick_rays_per_group_big(int S) {
  if (S > kMaxGroupPoints) return 0;
  int best = 0;
  double best_waste = 2.0;
  for (int g = 1; g * S <= kMaxGroupPoints; ++g) {
    const int pts = g * S;
    const int rounds = (pts + 127) / 128;
    const double waste = 1.0 - (double)pts / (rounds * 128);
    if (waste < best_waste - 1e-9 || (waste < best_waste + 1e-9 && g > best)) {
      best_waste = waste;
      best = g;
    }
  }
  return best;
}

template <bool kExport>
int configure_big() {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_render_forward_big_kernel<kExport>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tc_render_forward_big_kernel)");
    configured = true;
  }
  return FNR_OK;
}

}  // namespace

bool tc_big_supported(int S) { return S >= 1 && S <= kMaxGroupPoints; }

int launch_tc_render_forward_big(const KField& F, const KParams& P, const KRays& Rr, const KFieldOut& O, const KComposite& Cm, cudaStream_t st) {
  if (Rr.R == 0) return FNR_OK;
  if (int rc = configure_big<false>()) return rc;
  BigArgs a;
  memset(&a.E, 0, sizeof(a.E));
  a.F = F;
  a.P = P;
  a.Rr = Rr;
  a.O = O;
  a.Cm = Cm;
  a.rays_per_group = pick_rays_per_group_big(Rr.S);
  a.composite = Cm.rgb || Cm.accumulation || Cm.depth || Cm.depth_index || Cm.semantics || Cm.weights;
  const int groups = (Rr.R + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  tc_render_forward_big_kernel<false><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_big_kernel");
}

int launch_tc_export_big(const KField& F, const KParams& P, const KExport& E, cudaStream_t st) {
  if (E.B == 0) return FNR_OK;
  if (int rc = configure_big<true>()) return rc;
  BigArgs a;
  memset(&a, 0, sizeof(a));
  a.F = F;
  a.F.position_mode = FNR_POS_AABB;
  a.F.appearance_mode = FNR_APP_MEAN;
  a.P = P;
  a.E = E;
  a.rays_per_group = pick_rays_per_group_big(E.S);
  const int groups = (E.B + a.rays_per_group - 1) / a.rays_per_group;
  const int grid = groups < sm_count() ? groups : sm_count();
  tc_render_forward_big_kernel<true><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return check_launch("tc_render_forward_big_kernel<export>");
}

}  // namespace fnr

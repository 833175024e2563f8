// Per-group stage shared by the fused tcgen05 forward kernels (fnr_tc.cu, fnr_tc_big.cu): the group's per-sample
// density / rgb / logit sit in shared memory ([point][5] floats); a CTA either composites one ray per warp
// (fruit_nerf/fruit_nerf.py:325-348) or thresholds + compacts the export sets (fruit_nerf.py:251-269;
// export/exporter_utils.py:111-153).
#pragma once
#include "fnr_common.cuh"
#include "fnr_kernels.h"
#include "fnr_tc_common.cuh"

namespace fnr {
namespace tcx {

// one atomic per warp and output set: returns the claimed base row, rank_out = this lane's offset
__device__ __forceinline__ int warp_claim_rows(int* counter, bool pred, int lane, int& rank_out) {
  const unsigned m = __ballot_sync(kTcFullMask, pred);
  int basev = 0;
  if (m) {
    const int leader = __ffs(m) - 1;
    if (lane == leader) basev = atomicAdd(counter, __popc(m));
    basev = __shfl_sync(kTcFullMask, basev, leader);
  }
  rank_out = __popc(m & ((1u << lane) - 1u));
  return basev;
}

// UniformSamplerWithNoise: t = bins*far + (1-bins)*near (components/ray_samplers.py:89-94); bins are shared by all rays
// (eval mode) or per ray (stratified jitter of the training-mode module, ray_samplers.py:78-87)
__device__ __forceinline__ void export_interval(const KExport& E, int ray, int s, float& t0, float& t1) {
  const float* bins = E.bins + (size_t)ray * E.bins_ray_stride;
  const float b0 = __ldg(bins + s), b1 = __ldg(bins + s + 1);
  t0 = __fadd_rn(__fmul_rn(b0, E.far_plane), __fmul_rn(__fsub_rn(1.0f, b0), E.near_plane));
  t1 = __fadd_rn(__fmul_rn(b1, E.far_plane), __fmul_rn(__fsub_rn(1.0f, b1), E.near_plane));
}

template <int kCtaThreads>
__device__ __forceinline__ void group_export(const KExport& E, const KField& F, const float* s_samples, int ray0, int pts, int S) {
  const int tid = threadIdx.x, lane = tid & 31;
  // FruitModel.get_export_outputs + the selection body of sample_volume (fruit_nerf.py:251-269;
  // export/exporter_utils.py:111-153): three output sets, warp-aggregated row claims
    const size_t gbase = (size_t)ray0 * S;
  const int padded = (pts + 31) & ~31;
  for (int i = tid; i < padded; i += kCtaThreads) {
    const bool in = i < pts;
    const int ic = in ? i : pts - 1;
    const float* q = s_samples + 5 * ic;
    const size_t p = gbase + ic;
    float t0, t1;
    export_interval(E, ray0 + ic / S, ic % S, t0, t1);
    bool sel;
    Vec3 world;
    (void)field_position(E.origins + 3 * (size_t)(ray0 + ic / S), E.normal, t0, t1, FNR_POS_AABB, F.aabb, sel, &world);
    const float density = q[0], logit = q[4];
    const float sg = sigmoidf_(logit);
    const int label = (sg - E.label_thr > 0.f) ? 1 : 0;  // heaviside(sigmoid(logit) - thr, 0)
    if (in) {
      if (E.sample_density) E.sample_density[p] = density;
      if (E.sample_semantics) E.sample_semantics[p] = logit;
      if (E.semantics_colormap) E.semantics_colormap[p] = label;
      if (E.sample_rgb) {
        E.sample_rgb[3 * p] = q[1];
        E.sample_rgb[3 * p + 1] = q[2];
        E.sample_rgb[3 * p + 2] = q[3];
      }
      if (E.point_location) {
        E.point_location[3 * p] = world.x;
        E.point_location[3 * p + 1] = world.y;
        E.point_location[3 * p + 2] = world.z;
      }
    }
    const bool m_den = in && (density >= E.density_min);
    const bool m_sem = in && (logit >= E.logit_min);
    const bool m_lab = in && ((float)label >= 0.999f);
    const bool keep[3] = {m_lab && m_den, m_sem && m_den, m_den};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int rank;
      const int basev = warp_claim_rows(E.counts + k, keep[k], lane, rank);
      if (keep[k] && E.rows[k]) {
        const int orow = basev + rank;
        if (orow < E.capacity) {
          float* w = E.rows[k] + 7 * (size_t)orow;
          w[0] = world.x;
          w[1] = world.y;
          w[2] = world.z;
          w[3] = q[1];
          w[4] = q[2];
          w[5] = q[3];
          w[6] = (k == 2) ? sigmoidf_(density) : sg;
          if (E.keys[k]) E.keys[k][orow] = E.point_base + (uint64_t)p;
        }
      }
    }
  }

}

template <int kCtaThreads>
__device__ __forceinline__ void group_write_samples(const KFieldOut& O, const float* s_samples, int ray0, int pts, int S) {
  const int tid = threadIdx.x;
  const size_t gbase = (size_t)ray0 * S;
  for (int i = tid; i < pts; i += kCtaThreads) {
    const float* q = s_samples + 5 * i;
    if (O.sample_density) O.sample_density[gbase + i] = q[0];
    if (O.sample_semantics) O.sample_semantics[gbase + i] = q[4];
    if (O.sample_rgb) {
      O.sample_rgb[3 * (gbase + i)] = q[1];
      O.sample_rgb[3 * (gbase + i) + 1] = q[2];
      O.sample_rgb[3 * (gbase + i) + 2] = q[3];
    }
  }
}

template <int kCtaThreads>
__device__ __forceinline__ void group_composite(const KComposite& Cm, const KRays& Rr, const float* s_samples, int ray0, int rays_here, int S) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr unsigned kFullMask = kTcFullMask;
    for (int rl = warp; rl < rays_here; rl += kCtaThreads / 32) {
    const int r = ray0 + rl;
    const size_t base = (size_t)r * S;
    const float* sp = s_samples + 5 * (rl * S);
    float run_x = 0.f, run_w = 0.f, acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, sem = 0.f;
    int median = S;
    for (int c0 = 0; c0 < S; c0 += 32) {
      const int i = c0 + lane;
      const bool in = i < S;
      float x = 0.f, w = 0.f;
      if (in) x = (Rr.ends[base + i] - Rr.starts[base + i]) * sp[5 * i];
      const float incl = warp_incl_scan_f(x, lane);
      // exclusive prefix by shuffle, not `incl - x`: an infinite sigma*delta must give T = 1 in front of it (torch.cumsum semantics)
      float excl = __shfl_up_sync(kFullMask, incl, 1);
      if (lane == 0) excl = 0.f;
      if (in) {
        const float alpha = 1.0f - expf(-x);
        const float T = expf(-(run_x + excl));
        w = nan_to_num(alpha * T);
        if (Cm.weights) Cm.weights[base + i] = w;
        float c0r = sp[5 * i + 1], c0g = sp[5 * i + 2], c0b = sp[5 * i + 3];
        if (Cm.clamp_rgb) {
          c0r = nan_to_num(c0r);
          c0g = nan_to_num(c0g);
          c0b = nan_to_num(c0b);
        }
        cr += w * c0r;
        cg += w * c0g;
        cb += w * c0b;
        sem += w * sp[5 * i + 4];
        acc += w;
      }
      const float wincl = warp_incl_scan_f(w, lane);
      const unsigned m = __ballot_sync(kFullMask, in && (run_w + wincl >= 0.5f));
      if (m && median == S) median = c0 + (__ffs(m) - 1);
      run_x += __shfl_sync(kFullMask, incl, 31);
      run_w += __shfl_sync(kFullMask, wincl, 31);
    }
    acc = warp_sum_f(acc);
    cr = warp_sum_f(cr);
    cg = warp_sum_f(cg);
    cb = warp_sum_f(cb);
    sem = warp_sum_f(sem);
    if (lane == 0) {
      float lr = sp[5 * (S - 1) + 1], lg = sp[5 * (S - 1) + 2], lb = sp[5 * (S - 1) + 3];
      if (Cm.clamp_rgb) {
        lr = nan_to_num(lr);
        lg = nan_to_num(lg);
        lb = nan_to_num(lb);
      }
      float orr = cr + lr * (1.0f - acc), og = cg + lg * (1.0f - acc), ob = cb + lb * (1.0f - acc);
      if (Cm.clamp_rgb) {
        orr = fminf(fmaxf(orr, 0.f), 1.f);
        og = fminf(fmaxf(og, 0.f), 1.f);
        ob = fminf(fmaxf(ob, 0.f), 1.f);
      }
      if (Cm.rgb) {
        Cm.rgb[3 * r] = orr;
        Cm.rgb[3 * r + 1] = og;
        Cm.rgb[3 * r + 2] = ob;
      }
      if (Cm.accumulation) Cm.accumulation[r] = acc;
      if (Cm.semantics) Cm.semantics[r] = sem;
      const int mi = median < S - 1 ? median : S - 1;
      if (Cm.depth_index) Cm.depth_index[r] = mi;
      if (Cm.depth) Cm.depth[r] = (Rr.starts[base + mi] + Rr.ends[base + mi]) / 2;
    }
  }
}

}  // namespace tcx
}  // namespace fnr
